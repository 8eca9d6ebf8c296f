// Instantiates the CSR / COO kernels for ONE element type and provides its launchers.
// Compiled six times: -DPYGIM_T=int8_t -DPYGIM_SFX=i8, ... (see build.py).
#include "launch.h"
#include "spmm_coo.cuh"
#include "spmm_csr.cuh"

#ifndef PYGIM_T
#error "compile with -DPYGIM_T=<type> -DPYGIM_SFX=<suffix>"
#endif

namespace pygim {

#define PYGIM_CAT2(a, b) a##b
#define PYGIM_CAT(a, b) PYGIM_CAT2(a, b)

using T = PYGIM_T;

static inline int align_bytes(const void *p) { return (int)(reinterpret_cast<uintptr_t>(p) & 15); }

// 16-byte words when every row start is 16-byte aligned, single elements otherwise
static bool can_vectorize(const void *B, const void *C, long long ncols, long long ldb, long long ldc, long long ldp) {
    const long long s = (long long)sizeof(T);
    return align_bytes(B) == 0 && align_bytes(C) == 0 && (ncols * s) % 16 == 0 && (ldb * s) % 16 == 0 &&
           (ldc * s) % 16 == 0 && (ldp * s) % 16 == 0;
}

static int pow2_ceil(long long v) {
    int g = 1;
    while (g < v && g < 32) g <<= 1;
    return g;
}

// Tuning knobs (overridable at build time for experiments: -DPYGIM_CSR_UNROLL=.. etc.)
//   UNROLL     independent gathers per lane in flight before the first FMA
//   R          index entries per lane per batch (batch = 32*R nonzeros); R*G >= UNROLL keeps UNROLL usable
//   D          batches of the index stream prefetched ahead
//   MIN_BLOCKS resident 256-thread blocks per SM the register allocation must allow
#ifndef PYGIM_CSR_UNROLL
#define PYGIM_CSR_UNROLL 8
#endif
#ifndef PYGIM_CSR_MINBLOCKS
#define PYGIM_CSR_MINBLOCKS 4
#endif
#ifndef PYGIM_CSR_PREFETCH
#define PYGIM_CSR_PREFETCH 2
#endif
#ifndef PYGIM_NARROW_UNROLL
#define PYGIM_NARROW_UNROLL 4
#endif
#ifndef PYGIM_NARROW_MINBLOCKS
#define PYGIM_NARROW_MINBLOCKS 2
#endif
// 8/16-bit types carry E = 16/8 32-bit accumulators per lane: fewer gathers in flight and a larger register
// budget keep those instantiations spill-free.
template <int E, int G> struct CsrTune {
    static constexpr int UNROLL = (E >= 8) ? PYGIM_NARROW_UNROLL : PYGIM_CSR_UNROLL;
    static constexpr int R = (G >= UNROLL) ? 1 : ((UNROLL / G) > 4 ? 4 : (UNROLL / G));
    static constexpr int D = (R > 1) ? 1 : PYGIM_CSR_PREFETCH;
    static constexpr int MIN_BLOCKS =
        (E >= 8) ? PYGIM_NARROW_MINBLOCKS : ((sizeof(T) * E >= 16) ? PYGIM_CSR_MINBLOCKS : 4);
};
// Short-row graphs (mean degree < ~100: ogbn-products, citation graphs) are bound by the dependent chain of one
// row (index load -> gather -> shuffle tree -> store), not by gathers in flight: more resident warps win
// (measured on products-shape: H=16 1.66 -> 0.93 ms, H=32 1.93 -> 1.34 ms, H=64 2.92 -> 2.36 ms).
#ifndef PYGIM_SHORT_UNROLL
#define PYGIM_SHORT_UNROLL 2
#endif
#ifndef PYGIM_SHORT_MINBLOCKS
#define PYGIM_SHORT_MINBLOCKS 6
#endif
#ifndef PYGIM_SHORT_PREFETCH
#define PYGIM_SHORT_PREFETCH 1
#endif
template <int E, int G> struct CsrTuneShort {
    static constexpr int UNROLL = (E >= 8) ? PYGIM_NARROW_UNROLL : PYGIM_SHORT_UNROLL;
    static constexpr int R = (G >= UNROLL) ? 1 : ((UNROLL / G) > 4 ? 4 : (UNROLL / G));
    static constexpr int D = (R > 1) ? 1 : PYGIM_SHORT_PREFETCH;
    static constexpr int MIN_BLOCKS = (E >= 8) ? PYGIM_NARROW_MINBLOCKS : ((G >= 32) ? PYGIM_CSR_MINBLOCKS : PYGIM_SHORT_MINBLOCKS);
};
// Streamed row tickets (short_rows == 2): 4 gathers of a run in flight, batches of 32 prefetched 2 ahead
#ifndef PYGIM_STREAM_MINBLOCKS
#define PYGIM_STREAM_MINBLOCKS 4
#endif
template <int E, int G> struct CsrTuneStream {
    static constexpr int UNROLL = 4;
    static constexpr int R = 1;
    static constexpr int D = 2;
    static constexpr int MIN_BLOCKS = PYGIM_STREAM_MINBLOCKS;
};
// COO carries a third index stream and the run walker: one block less per SM than CSR keeps it spill-free
template <int E, int G> struct CooTune {
    static constexpr int UNROLL = CsrTune<E, G>::UNROLL;
    static constexpr int R = CsrTune<E, G>::R;
    static constexpr int D = CsrTune<E, G>::D;
    static constexpr int MIN_BLOCKS = CsrTune<E, G>::MIN_BLOCKS > 3 ? 3 : CsrTune<E, G>::MIN_BLOCKS;
};


template <int E, int G, bool UNIT, typename Tune, bool STREAM = false>
static cudaError_t launch_csr_t(CsrArgs<T> a, const CsrLaunch &l, int64_t *launches) {
    auto kernel = csr_spmm_kernel<T, E, G, Tune::UNROLL, Tune::MIN_BLOCKS, Tune::R, Tune::D, UNIT, STREAM>;
    static int blocks_per_sm = 0;   // per instantiation
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kCsrThreads, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    a.col_chunks = (a.nvec + G - 1) / G;
    const unsigned long long total =
        (unsigned long long)a.col_chunks * ((unsigned long long)a.n_seg + (unsigned long long)a.n_row_tickets);
    const unsigned long long warps_needed = total;
    unsigned long long blocks = (unsigned long long)blocks_per_sm * (l.sm_count > 0 ? l.sm_count : 148);
    const unsigned long long blocks_needed = (warps_needed + (kCsrThreads / 32) - 1) / (kCsrThreads / 32);
    if (blocks > blocks_needed) blocks = blocks_needed;
    a.ticket = l.ticket;
    a.n_warps = (unsigned)(blocks * (kCsrThreads / 32));
    kernel<<<(unsigned)blocks, kCsrThreads, 0, l.stream>>>(a);
    ++*launches;
    return cudaGetLastError();
}

template <int E, int G, bool UNIT>
static cudaError_t launch_csr_g(const CsrArgs<T> &a, const CsrLaunch &l, int64_t *launches) {
    // the 16-byte-word instantiations of 32/64-bit types come in two register budgets (see CsrTuneShort)
    // (rows of 512 bytes and more - G == 32 - gain nothing from either: measured 5.11 vs 4.97 ms on products-shape)
    if constexpr (E < 8 && sizeof(T) * E >= 16 && G < 32) {
        if (l.short_rows == 2) return launch_csr_t<E, G, UNIT, CsrTuneStream<E, G>, true>(a, l, launches);   // streamed
        if (l.short_rows) return launch_csr_t<E, G, UNIT, CsrTuneShort<E, G>>(a, l, launches);
    }
    return launch_csr_t<E, G, UNIT, CsrTune<E, G>>(a, l, launches);
}

template <int E> static cudaError_t launch_csr_e(const CsrLaunch &l, int64_t *launches) {
    CsrArgs<T> a;
    a.rowptr = l.rowptr;
    a.colind = l.colind;
    a.val = static_cast<const T *>(l.val);
    a.B = static_cast<const T *>(l.B);
    a.C = static_cast<T *>(l.C);
    a.partial = static_cast<T *>(l.partial);
    a.segs = l.segs;
    a.long_rows = l.long_rows;
    a.long_seg_ptr = l.long_seg_ptr;
    a.seg_count = l.seg_count;
    a.n_long = l.n_long;
    a.n_seg = l.n_seg;
    a.nrows = l.nrows;
    a.seg_len = l.seg_len;
    a.rows_per_ticket = l.rows_per_ticket < 1 ? 1 : (l.rows_per_ticket > 31 ? 31 : l.rows_per_ticket);
    a.n_row_tickets = (l.nrows + a.rows_per_ticket - 1) / a.rows_per_ticket;
    a.nvec = (int)(l.ncols / E);
    a.ldb = l.ldb;
    a.ldb_bytes = (unsigned)(l.ldb * (long long)sizeof(T));
    a.ldc = l.ldc;
    a.ldp = l.ldp;
    a.accumulate = l.accumulate;
    a.n_peers = l.n_peers;
    a.mc = static_cast<T *>(l.mc);
    for (int p = 0; p < kMaxPeers; ++p) a.peers[p] = p < l.n_peers ? static_cast<T *>(l.peers[p]) : nullptr;
    const long long items = (long long)l.n_seg + l.nrows;
    if (items == 0 || a.nvec == 0) return cudaSuccess;
    cudaError_t err;
#define PYGIM_CSR_CASE(GV) \
    case GV: err = l.unit_values ? launch_csr_g<E, GV, true>(a, l, launches) : launch_csr_g<E, GV, false>(a, l, launches); break
    switch (pow2_ceil(a.nvec)) {
        PYGIM_CSR_CASE(1);
        PYGIM_CSR_CASE(2);
        PYGIM_CSR_CASE(4);
        PYGIM_CSR_CASE(8);
        PYGIM_CSR_CASE(16);
        default: err = l.unit_values ? launch_csr_g<E, 32, true>(a, l, launches) : launch_csr_g<E, 32, false>(a, l, launches); break;
    }
#undef PYGIM_CSR_CASE
    return err;
}

// flag[0] is cleared when any stored value differs from one (plan-time check for the unit-value fast path)
__global__ void all_ones_kernel(const T *val, long long n, int *flag) {
    bool ok = true;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        ok = ok && (val[i] == (T)1);
    if (!ok) *flag = 0;
}

cudaError_t PYGIM_CAT(check_all_ones_, PYGIM_SFX)(const void *val, long long n, int *d_flag, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    all_ones_kernel<<<1184, 256, 0, stream>>>(static_cast<const T *>(val), n, d_flag);
    return cudaGetLastError();
}

cudaError_t PYGIM_CAT(launch_csr_, PYGIM_SFX)(const CsrLaunch &l, int64_t *launches) {
    if (can_vectorize(l.B, l.C, l.ncols, l.ldb, l.ldc, l.ldp)) return launch_csr_e<16 / (int)sizeof(T)>(l, launches);
    return launch_csr_e<1>(l, launches);
}

template <int E, int G, bool UNIT>
static cudaError_t launch_coo_g(CooArgs<T> a, const CooLaunch &l, int64_t *launches) {
    auto kernel = coo_spmm_kernel<T, E, G, CooTune<E, G>::UNROLL, CooTune<E, G>::MIN_BLOCKS, CooTune<E, G>::R,
                                  CooTune<E, G>::D, UNIT>;
    static int blocks_per_sm = 0;   // per instantiation
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kCooThreads, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    a.col_chunks = (a.nvec + G - 1) / G;
    const unsigned long long total = (unsigned long long)a.col_chunks * (unsigned long long)a.n_chunks;
    unsigned long long blocks = (unsigned long long)blocks_per_sm * (l.sm_count > 0 ? l.sm_count : 148);
    const unsigned long long blocks_needed = (total + (kCooThreads / 32) - 1) / (kCooThreads / 32);
    if (blocks > blocks_needed) blocks = blocks_needed;
    a.ticket = l.ticket;
    a.n_warps = (unsigned)(blocks * (kCooThreads / 32));
    kernel<<<(unsigned)blocks, kCooThreads, 0, l.stream>>>(a);
    ++*launches;
    return cudaGetLastError();
}

template <int E> static cudaError_t launch_coo_e(const CooLaunch &l, int64_t *launches) {
    CooArgs<T> a;
    a.rowind = l.rowind;
    a.colind = l.colind;
    a.val = static_cast<const T *>(l.val);
    a.B = static_cast<const T *>(l.B);
    a.C = static_cast<T *>(l.C);
    a.nnz = l.nnz;
    a.ldb = l.ldb;
    a.ldb_bytes = (unsigned)(l.ldb * (long long)sizeof(T));
    a.ldc = l.ldc;
    a.nvec = (int)(l.ncols / E);
    a.accumulate = l.accumulate;
    cudaError_t err;
    if (!l.accumulate && l.nrows > 0 && l.ncols > 0) {
        // the reference hands the kernels a torch::zeros result (pytorch_api.cpp:357-358)
        err = cudaMemset2DAsync(l.C, (size_t)l.ldc * sizeof(T), 0, (size_t)l.ncols * sizeof(T), (size_t)l.nrows,
                                l.stream);
        if (err != cudaSuccess) return err;
        ++*launches;
    }
    if (l.nnz == 0 || a.nvec == 0) return cudaSuccess;
    // exact equal-nnz chunks (BLNC_NNZ): ~8 chunks per resident warp, 256..4096 nonzeros each, multiple of 32
    long long chunk = l.chunk_nnz;
    if (chunk <= 0) {
        const long long want = l.nnz / ((long long)(l.n_warp_slots > 0 ? l.n_warp_slots : 9472) * 8);
        chunk = 256;
        while (chunk < want && chunk < 4096) chunk <<= 1;
    }
    chunk = (chunk + 31) / 32 * 32;
    a.chunk_nnz = (int)chunk;
    a.n_chunks = (l.nnz + chunk - 1) / chunk;
#define PYGIM_COO_CASE(GV) \
    case GV: return l.unit_values ? launch_coo_g<E, GV, true>(a, l, launches) : launch_coo_g<E, GV, false>(a, l, launches)
    switch (pow2_ceil(a.nvec)) {
        PYGIM_COO_CASE(1);
        PYGIM_COO_CASE(2);
        PYGIM_COO_CASE(4);
        PYGIM_COO_CASE(8);
        PYGIM_COO_CASE(16);
        default: return l.unit_values ? launch_coo_g<E, 32, true>(a, l, launches) : launch_coo_g<E, 32, false>(a, l, launches);
    }
#undef PYGIM_COO_CASE
}

cudaError_t PYGIM_CAT(launch_coo_, PYGIM_SFX)(const CooLaunch &l, int64_t *launches) {
    if (can_vectorize(l.B, l.C, l.ncols, l.ldb, l.ldc, 16)) return launch_coo_e<16 / (int)sizeof(T)>(l, launches);
    return launch_coo_e<1>(l, launches);
}

}  // namespace pygim
