// Instantiates the CSR / COO kernels for ONE element type and provides its launchers.
// Compiled six times: -DPYGIM_T=int8_t -DPYGIM_SFX=i8, ... (see build.py).
#include "launch.h"
#include "spmm_coo.cuh"
#include "spmm_csr.cuh"
#include "spmm_csr_hc.cuh"

#ifndef PYGIM_T
#error "compile with -DPYGIM_T=<type> -DPYGIM_SFX=<suffix>"
#endif

namespace pygim {

#define PYGIM_CAT2(a, b) a##b
#define PYGIM_CAT(a, b) PYGIM_CAT2(a, b)

using T = PYGIM_T;

// Register budgets / gathers in flight per instantiation family (overridable at build time for experiments).
//   NV         index vectors (4 nonzeros each) per lane in flight: 4*NV independent gathers before the first FMA
//   THREADS    upper bound of the block size the launcher may pick; 65536 / THREADS is the register budget
//   MIN_BLOCKS resident blocks of THREADS threads per SM the register allocation must allow
// What decides is whether ptxas can keep ALL 4*NV gathers in flight: with the scheduler state of the persistent
// kernel live across the loop, 64 registers only fit ~3 (it consumes each gather right after issuing it); measured
// on Reddit-shape (sweep GFLOP/s): 64 regs NV=2 8500, 85 regs NV=2 8920, 128 regs NV=3 9150, 128 regs NV=4 9290.
#ifndef PYGIM_CSR_NV
#define PYGIM_CSR_NV 4      // 0 = the shuffle-delivered index stream of round 1 (csr_accumulate_shfl)
#endif
#ifndef PYGIM_CSR_THREADS
#define PYGIM_CSR_THREADS 512
#endif
#ifndef PYGIM_CSR_NV_WEIGHTED
#define PYGIM_CSR_NV_WEIGHTED 2
#endif
// DEEP (default for long rows): 128 registers, 16 resident warps per SM, 16 gathers per lane in flight.
// 8/16-bit types carry E = 16/8 32-bit accumulators per lane: four gathers.  Weighted (non-unit) kernels also
// hold the values of the nonzeros in flight: eight gathers.
template <int E, bool UNIT> struct CsrTune {
    static constexpr int NV = PYGIM_CSR_NV == 0 ? 0 : (E >= 8 ? 1 : (!UNIT ? PYGIM_CSR_NV_WEIGHTED : PYGIM_CSR_NV));
    static constexpr int THREADS = PYGIM_CSR_THREADS;
    static constexpr int MIN_BLOCKS = 1;
};
// LIGHT (short_rows == 3, default for short rows): 64 registers, 32 resident warps per SM.  Short-row graphs
// (mean degree < ~100: ogbn-products, citation graphs) are bound by the dependent chain of one row (index load ->
// gather -> shuffle tree -> store), not by gathers in flight: more resident warps win (products-shape sweep:
// 2656 GFLOP/s vs 2314 with the deep budget).
template <int E, bool UNIT> struct CsrTuneLight {
    static constexpr int NV = PYGIM_CSR_NV == 0 ? 0 : ((E >= 8 || !UNIT) ? 1 : 2);
    static constexpr int THREADS = (E >= 8) ? 512 : 1024;
    static constexpr int MIN_BLOCKS = 1;
};
// HIGH OCCUPANCY (short_rows == 1): 40 registers, 48 resident warps
template <int E, bool UNIT> struct CsrTuneShort {
    static constexpr int NV = PYGIM_CSR_NV == 0 ? 0 : 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = (E >= 8) ? 2 : 6;
};
// Streamed row items (short_rows == 2): the row path is csr_stream_rows; the segment path keeps four gathers in
// flight so the two code paths share one 64-register budget
template <int E, bool UNIT> struct CsrTuneStream {
    static constexpr int NV = PYGIM_CSR_NV == 0 ? 0 : 1;
    static constexpr int THREADS = (E >= 8) ? 512 : 1024;
    static constexpr int MIN_BLOCKS = 1;
};
// COO carries a third index stream and the run walker: one block less per SM than CSR keeps it spill-free
#ifndef PYGIM_CSR_UNROLL
#define PYGIM_CSR_UNROLL 8
#endif
template <int E, int G> struct CooTune {
    static constexpr int UNROLL = (E >= 8) ? 4 : PYGIM_CSR_UNROLL;
    static constexpr int R = (G >= UNROLL) ? 1 : ((UNROLL / G) > 4 ? 4 : (UNROLL / G));
    static constexpr int D = (R > 1) ? 1 : 2;
    static constexpr int MIN_BLOCKS = (E >= 8) ? 2 : 3;
};

template <int E, int G, bool UNIT, typename Tune, int STREAM = 0>
static cudaError_t launch_csr_t(CsrArgs<T> a, const CsrLaunch &l, int64_t *launches) {
    auto kernel = csr_spmm_kernel<T, E, G, Tune::NV, Tune::THREADS, Tune::MIN_BLOCKS, UNIT, STREAM>;
    int threads = l.cta_threads > 0 ? l.cta_threads : 256;
    threads = (threads + 31) / 32 * 32;
    if (threads > Tune::THREADS) threads = Tune::THREADS;
    static int blocks_per_sm[33] = {0};   // per instantiation, indexed by warps per block
    int &bps = blocks_per_sm[threads / 32];
    if (bps == 0) {
        // no shared memory is used: give the whole unified array to the L1 (the gathers' only on-SM reuse)
        (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        (void)cudaGetLastError();
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, threads, 0);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
    }
    a.col_chunks = (a.nvec + G - 1) / G;
    const int warps_per_block = threads / 32;
    // no more warps than items (small graphs: a short grid drains and leaves faster)
    const unsigned long long items = (unsigned long long)l.n_items * (unsigned long long)a.col_chunks;
    unsigned long long blocks = (unsigned long long)bps * (l.sm_count > 0 ? l.sm_count : 148);
    const unsigned long long blocks_needed = (items + warps_per_block - 1) / warps_per_block;
    if (blocks > blocks_needed) blocks = blocks_needed;
    if (blocks < 1) blocks = 1;
    a.n_warps = (unsigned)(blocks * warps_per_block);
    kernel<<<(unsigned)blocks, threads, 0, l.stream>>>(a);
    ++*launches;
    return cudaGetLastError();
}

template <int E, int G, bool UNIT>
static cudaError_t launch_csr_g(const CsrArgs<T> &a, const CsrLaunch &l, int64_t *launches) {
    // the 16-byte-word instantiations of 32/64-bit types come in two register budgets (see CsrTuneShort)
    // (rows of 512 bytes and more - G == 32 - gain nothing from either: measured 5.11 vs 4.97 ms on products-shape)
    if constexpr (E < 8 && sizeof(T) * E >= 16 && G < 32) {
        if (l.short_rows == 2) return launch_csr_t<E, G, UNIT, CsrTuneStream<E, UNIT>, 1>(a, l, launches);   // streamed
        if (l.short_rows == 1) return launch_csr_t<E, G, UNIT, CsrTuneShort<E, UNIT>>(a, l, launches);
    }
    if (l.short_rows == 3) return launch_csr_t<E, G, UNIT, CsrTuneLight<E, UNIT>>(a, l, launches);
    if (l.short_rows == 4) {
        // tiny rows (<= kTinyRow nonzeros) by a matrix-wide grid of lane groups, the rest by the persistent kernel
        // (which skips the tiny rows); the launches write disjoint rows
        const int col_chunks = (a.nvec + G - 1) / G;
        if (a.nrows > 0) {
            const unsigned bx = (unsigned)((a.nrows + (256 / G) - 1) / (256 / G));
            csr_tiny_rows_kernel<T, E, G, UNIT><<<dim3(bx, (unsigned)col_chunks), 256, 0, l.stream>>>(a);
            ++*launches;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
        // the pieces run on the DEEP register budget (16 warps per SM, 16 gathers in flight): measured on arxiv-shape,
        // H = 16 / 32 / 64 / 128, whole SpMM as a CUDA-graph replay - deep 32 / 37 / 54 / 104 us, light (32 warps) 39 / 43 /
        // 59 / 120 us, high occupancy (48 warps) 46 / 55 / 73 / 120 us: fewer warps on the one ticket counter, deeper rounds
        return launch_csr_t<E, G, UNIT, CsrTune<E, UNIT>, 2>(a, l, launches);
    }
    return launch_csr_t<E, G, UNIT, CsrTune<E, UNIT>>(a, l, launches);
}

// hot/cold kernel: one block per SM, dynamic shared memory = the tile
template <int E, int G, bool UNIT>
static cudaError_t launch_csr_hc(const CsrArgs<T> &a0, const CsrLaunch &l, int64_t *launches) {
#ifndef PYGIM_HC_THREADS
#define PYGIM_HC_THREADS 512
#endif
    constexpr int THREADS = PYGIM_HC_THREADS;
    constexpr int NV = (E >= 8) ? 1 : (THREADS > 768 ? 1 : (THREADS > 512 ? 2 : (UNIT ? 4 : 2)));
    auto kernel = csr_hc_kernel<T, E, G, NV, THREADS, UNIT>;
    const size_t smem = (size_t)l.hot_k * G * 16;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        smem_set = smem;
    }
    HcArgs<T> h;
    h.c = a0;
    h.c.col_chunks = (a0.nvec + G - 1) / G;
    h.hot_cols = l.hot_cols;
    h.hot_cnt = l.hot_cnt;
    h.hot_k = l.hot_k;
    h.n_seg_super = l.n_seg_super;
    int threads = THREADS;
    const int blocks = l.sm_count > 0 ? l.sm_count : 148;
    h.c.n_warps = (unsigned)(blocks * (threads / 32));
    kernel<<<blocks, threads, smem, l.stream>>>(h);
    ++*launches;
    return cudaGetLastError();
}

template <int E> static cudaError_t launch_csr_e(const CsrLaunch &l, int64_t *launches) {
    CsrArgs<T> a;
    a.rowptr = l.rowptr;
    a.colind = l.colind;
    a.val = static_cast<const T *>(l.val);
    a.B = static_cast<const T *>(l.B);
    a.C = l.C;
    a.partial = static_cast<T *>(l.partial);
    a.segs = l.segs;
    a.supers = l.supers;
    a.long_rows = l.long_rows;
    a.long_seg_ptr = l.long_seg_ptr;
    a.super_cnt = l.super_cnt;
    a.seg_count = l.seg_count;
    a.warps_out = l.warps_out;
    a.n_super = l.n_super;
    a.n_long = l.n_long;
    a.n_seg = l.n_seg;
    a.nrows = l.nrows;
    a.seg_len = l.seg_len;
    a.nvec = (int)(l.ncols / E);
    a.nnz_total = l.nnz_total;
    {   // vector index loads need colind and val misaligned by the same number of ELEMENTS (mod 4)
        const int mc = (int)((reinterpret_cast<uintptr_t>(l.colind) >> 2) & 3);
        const int mv = (int)((reinterpret_cast<uintptr_t>(l.val) / sizeof(T)) & 3);
        const bool ok = (reinterpret_cast<uintptr_t>(l.colind) & 3) == 0 && (l.unit_values || mc == mv);
        a.idx_mis = ok ? mc : 4;
    }
    a.ldb = l.ldb;
    a.ldb_bytes = (unsigned)(l.ldb * (long long)sizeof(T));
    a.ldc = l.ldc;
    a.ldp = l.ldp;
    a.accumulate = l.accumulate;
    a.epi.row_map = l.epi.row_map;
    a.epi.scale = l.epi.scale;
    a.epi.residual = l.epi.residual;
    a.epi.ld_res = l.epi.ld_res;
    a.epi.coeff = l.epi.coeff;
    a.epi.mc = l.epi.mc;
    a.epi.peer_mask = l.epi.peer_mask;
    a.epi.n_peers = l.epi.n_peers;
    a.epi.my_rank = l.epi.my_rank;
    a.epi.epoch = l.epi.epoch;
    for (int p = 0; p < kMaxPeers; ++p) {
        a.epi.peers[p] = p < l.epi.n_peers ? l.epi.peers[p] : nullptr;
        a.epi.flags[p] = p < l.epi.n_peers ? l.epi.flags[p] : nullptr;
    }
    if ((l.n_items == 0 && l.short_rows != 4) || a.nvec == 0) return cudaSuccess;   // family 4: the tiny-row launch has no items
    cudaError_t err;
    if (l.hot_k > 0) {
        if constexpr (sizeof(T) * E == 16) {
            const int g = csr_lanes(a.nvec, l.max_g < 8 ? l.max_g : 8);
            const bool u = l.unit_values != 0;
            if (g >= 8) return u ? launch_csr_hc<E, 8, true>(a, l, launches) : launch_csr_hc<E, 8, false>(a, l, launches);
            if (g == 4) return u ? launch_csr_hc<E, 4, true>(a, l, launches) : launch_csr_hc<E, 4, false>(a, l, launches);
        }
        return cudaErrorInvalidValue;     // hot/cold plans need 16-byte words and dense rows of at least 64 bytes
    }
#define PYGIM_CSR_CASE(GV) \
    case GV: err = l.unit_values ? launch_csr_g<E, GV, true>(a, l, launches) : launch_csr_g<E, GV, false>(a, l, launches); break
    switch (csr_lanes(a.nvec, l.max_g)) {
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
    const size_t out_elem = (l.epi.scale || l.epi.residual) ? sizeof(float) : sizeof(T);
    if (csr_can_vectorize(sizeof(T), l.B, l.C, out_elem, l.ncols, l.ldb, l.ldc, l.ldp))
        return launch_csr_e<16 / (int)sizeof(T)>(l, launches);
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
    a.all_atomic = l.all_atomic;
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
    switch (csr_lanes(a.nvec, 32)) {
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
    if (csr_can_vectorize(sizeof(T), l.B, l.C, sizeof(T), l.ncols, l.ldb, l.ldc, 16))
        return launch_coo_e<16 / (int)sizeof(T)>(l, launches);
    return launch_coo_e<1>(l, launches);
}

}  // namespace pygim
