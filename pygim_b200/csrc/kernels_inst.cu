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
#define PYGIM_CSR_MINBLOCKS 3
#endif
#ifndef PYGIM_CSR_PREFETCH
#define PYGIM_CSR_PREFETCH 2
#endif
template <int E, int G> struct CsrTune {
    static constexpr int UNROLL = PYGIM_CSR_UNROLL;
    static constexpr int R = (G >= UNROLL) ? 1 : ((UNROLL / G) > 4 ? 4 : (UNROLL / G));
    static constexpr int D = (R > 1) ? 1 : PYGIM_CSR_PREFETCH;
    static constexpr int MIN_BLOCKS = (sizeof(T) * E >= 16) ? PYGIM_CSR_MINBLOCKS : 4;
};

template <int E, int G> static cudaError_t launch_csr_g(CsrArgs<T> a, const CsrLaunch &l, int64_t *launches) {
    auto kernel = csr_spmm_kernel<T, E, G, CsrTune<E, G>::UNROLL, CsrTune<E, G>::MIN_BLOCKS, CsrTune<E, G>::R,
                                  CsrTune<E, G>::D>;
    static int blocks_per_sm = 0;   // per instantiation
    if (blocks_per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kCsrThreads, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    a.col_chunks = (a.nvec + G - 1) / G;
    const unsigned long long total = (unsigned long long)a.col_chunks * ((unsigned long long)a.n_seg + a.nrows);
    const unsigned long long warps_needed = total;
    unsigned long long blocks = (unsigned long long)blocks_per_sm * (l.sm_count > 0 ? l.sm_count : 148);
    const unsigned long long blocks_needed = (warps_needed + (kCsrThreads / 32) - 1) / (kCsrThreads / 32);
    if (blocks > blocks_needed) blocks = blocks_needed;
    a.ticket = l.ticket;
    a.ticket_base = *l.ticket_base;
    kernel<<<(unsigned)blocks, kCsrThreads, 0, l.stream>>>(a);
    // every warp draws tickets until it sees one past the end: total + (#warps) tickets per launch
    *l.ticket_base += total + blocks * (kCsrThreads / 32);
    ++*launches;
    return cudaGetLastError();
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
    a.n_seg = l.n_seg;
    a.nrows = l.nrows;
    a.seg_len = l.seg_len;
    a.nvec = (int)(l.ncols / E);
    a.ldb = l.ldb;
    a.ldc = l.ldc;
    a.ldp = l.ldp;
    a.accumulate = l.accumulate;
    const long long items = (long long)l.n_seg + l.nrows;
    if (items == 0 || a.nvec == 0) return cudaSuccess;
    cudaError_t err;
    switch (pow2_ceil(a.nvec)) {
        case 1: err = launch_csr_g<E, 1>(a, l, launches); break;
        case 2: err = launch_csr_g<E, 2>(a, l, launches); break;
        case 4: err = launch_csr_g<E, 4>(a, l, launches); break;
        case 8: err = launch_csr_g<E, 8>(a, l, launches); break;
        case 16: err = launch_csr_g<E, 16>(a, l, launches); break;
        default: err = launch_csr_g<E, 32>(a, l, launches); break;
    }
    if (err != cudaSuccess) return err;
    if (l.n_long > 0) {
        FixupArgs<T> f;
        f.partial = a.partial;
        f.C = a.C;
        f.long_rows = l.long_rows;
        f.long_seg_ptr = l.long_seg_ptr;
        f.ldp = l.ldp;
        f.ldc = l.ldc;
        f.ncols = (int)l.ncols;
        f.accumulate = l.accumulate;
        const int threads = l.ncols >= 128 ? 128 : (l.ncols > 32 ? 64 : 32);
        csr_fixup_kernel<T><<<l.n_long, threads, 0, l.stream>>>(f);
        ++*launches;
        err = cudaGetLastError();
    }
    return err;
}

cudaError_t PYGIM_CAT(launch_csr_, PYGIM_SFX)(const CsrLaunch &l, int64_t *launches) {
    if (can_vectorize(l.B, l.C, l.ncols, l.ldb, l.ldc, l.ldp)) return launch_csr_e<16 / (int)sizeof(T)>(l, launches);
    return launch_csr_e<1>(l, launches);
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
    const int G = pow2_ceil(a.nvec);
    const int P = 32 / G;
    // exact equal-nnz chunks (BLNC_NNZ): aim at ~8 chunks per resident warp, 64..2048 nonzeros each
    long long chunk = l.chunk_nnz;
    if (chunk <= 0) {
        chunk = l.nnz / ((long long)(l.n_warp_slots > 0 ? l.n_warp_slots : 9472) * 8);
        if (chunk < 64) chunk = 64;
        if (chunk > 2048) chunk = 2048;
    }
    long long sub = (chunk + P - 1) / P;
    sub = (sub + G - 1) / G * G;
    a.sub_nnz = (int)sub;
    const long long per_warp = sub * P;
    const long long warps = (l.nnz + per_warp - 1) / per_warp;
    dim3 grid((unsigned)((warps + kCooWarpsPerBlock - 1) / kCooWarpsPerBlock), (unsigned)((a.nvec + G - 1) / G));
    dim3 block(kCooWarpsPerBlock * 32);
    switch (G) {
        case 1: coo_spmm_kernel<T, E, 1><<<grid, block, 0, l.stream>>>(a); break;
        case 2: coo_spmm_kernel<T, E, 2><<<grid, block, 0, l.stream>>>(a); break;
        case 4: coo_spmm_kernel<T, E, 4><<<grid, block, 0, l.stream>>>(a); break;
        case 8: coo_spmm_kernel<T, E, 8><<<grid, block, 0, l.stream>>>(a); break;
        case 16: coo_spmm_kernel<T, E, 16><<<grid, block, 0, l.stream>>>(a); break;
        default: coo_spmm_kernel<T, E, 32><<<grid, block, 0, l.stream>>>(a); break;
    }
    ++*launches;
    return cudaGetLastError();
}

cudaError_t PYGIM_CAT(launch_coo_, PYGIM_SFX)(const CooLaunch &l, int64_t *launches) {
    if (can_vectorize(l.B, l.C, l.ncols, l.ldb, l.ldc, 16)) return launch_coo_e<16 / (int)sizeof(T)>(l, launches);
    return launch_coo_e<1>(l, launches);
}

}  // namespace pygim
