// COO SpMM / SpMV for sm_100a: C[row[e], :] += val[e] * B[col[e], :] over a row-major sorted stream.
//
// Replaces the DPU kernels spmm_default/dpu_kernels/spmm_mul_coo_dpu.c:49-400 and
// spmv_sparseP/dpu_kernels/spmv_mul_coo_dpu.c:41-693.  Like the reference's BLNC_NNZ /
// BLNC_TSKLT_NNZ policy (spmm_default/spmm_mul_coo.c:124-165, support/partition.c:231-261) the
// nonzero stream is cut into EXACT equal-nnz chunks, so a row may straddle two chunks.  The
// reference resolves those shared rows with a mutex (CG_LOCK) or a per-tasklet spill that one
// tasklet adds up afterwards (LOCKFREE / LOCKFREEV2, spmm_mul_coo_dpu.c:166-390) plus a host-side
// add of the DPU boundary rows (spmm_mul_coo.c:478-489).  Here:
//
//  * a warp owns chunk_nnz consecutive nonzeros, split into P = 32/G contiguous sub-chunks, one
//    per group of G lanes (G lanes x 16 bytes cover one dense row);
//  * each group streams its sub-chunk - G (row, col, val) triples per coalesced load, handed
//    round the group with width-G shuffles - and runs a segmented reduction in registers:
//    the accumulator is flushed whenever the row index changes;
//  * rows that lie wholly inside one sub-chunk are written with plain stores; only a sub-chunk's
//    first/last row, and only if the neighbouring nonzero really has the same row, is combined
//    with atomics (integer atomics are exact; float atomics commute up to rounding - see DESIGN.md).
//
// C is zero-filled by the caller before the launch (the reference's torch::zeros,
// pytorch_api.cpp:357-358) unless accumulate is set.
#pragma once
#include "vec.cuh"

namespace pygim {

template <typename T> struct CooArgs {
    const int *rowind;
    const int *colind;
    const T *val;
    const T *B;
    T *C;
    long long nnz;
    long long ldb, ldc;
    int sub_nnz;       // nonzeros per sub-chunk (multiple of G); a warp covers P * sub_nnz
    int nvec;          // words (of E elements) per dense row
    int accumulate;
};

constexpr int kCooWarpsPerBlock = 8;

template <typename T> __device__ __forceinline__ void atomic_add_elem(T *p, typename Arith<T>::Acc v) {
    if constexpr (std::is_same<T, float>::value || std::is_same<T, double>::value) {
        atomicAdd(p, (T)v);
    } else if constexpr (sizeof(T) == 4) {
        atomicAdd(reinterpret_cast<unsigned int *>(p), (unsigned int)v);
    } else if constexpr (sizeof(T) == 8) {
        atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
    } else {
        // 8/16-bit: compare-and-swap on the containing aligned 32-bit word, wrapping inside the field
        const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
        unsigned int *word = reinterpret_cast<unsigned int *>(addr & ~(uintptr_t)3);
        const unsigned int shift = (unsigned int)(addr & 3) * 8;
        const unsigned int mask = (sizeof(T) == 1 ? 0xffu : 0xffffu) << shift;
        unsigned int old = *word, assumed;
        do {
            assumed = old;
            const unsigned int field = (((assumed & mask) >> shift) + (unsigned int)v) << shift;
            old = atomicCAS(word, assumed, (assumed & ~mask) | (field & mask));
        } while (old != assumed);
    }
}

template <typename T, int E>
__device__ __forceinline__ void coo_flush(T *dst, typename Arith<T>::Acc (&acc)[E], bool exclusive, bool accumulate) {
    if (exclusive) {
        if (accumulate) add_old<T, E>(acc, ld_plain<T, E>(dst));
        st_plain<T, E>(dst, narrow<T, E>(acc));
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) atomic_add_elem<T>(dst + k, acc[k]);
    }
}

// grid.x = ceil(n_chunks / kCooWarpsPerBlock); grid.y = column chunks of G words (only > 1 when G == 32)
template <typename T, int E, int G>
__global__ void __launch_bounds__(kCooWarpsPerBlock * 32) coo_spmm_kernel(const CooArgs<T> a) {
    using Acc = typename Arith<T>::Acc;
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    constexpr int UNROLL = (G < 8) ? G : 8;
    constexpr unsigned FULL = 0xffffffffu;

    const long long warp = (long long)blockIdx.x * kCooWarpsPerBlock + (threadIdx.x >> 5);
    const long long chunk_start = warp * (long long)P * a.sub_nnz;
    if (chunk_start >= a.nnz) return;   // whole warp leaves together
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, l = lane % G;
    const int vec = blockIdx.y * G + l;
    const bool active = vec < a.nvec;
    const T *Bcol = a.B + (long long)vec * E;
    T *Ccol = a.C + (long long)vec * E;

    long long s_begin = chunk_start + (long long)sub * a.sub_nnz;
    long long s_end = s_begin + a.sub_nnz;
    if (s_begin > a.nnz) s_begin = a.nnz;
    if (s_end > a.nnz) s_end = a.nnz;
    const bool nonempty = s_begin < s_end;

    // does the neighbouring nonzero belong to the same row as our first / last one?
    int first_row = -1;
    bool shared_head = false, shared_tail = false;
    if (nonempty) {
        first_row = a.rowind[s_begin];
        shared_head = s_begin > 0 && a.rowind[s_begin - 1] == first_row;
        shared_tail = s_end < a.nnz && a.rowind[s_end] == a.rowind[s_end - 1];
    }

    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    int cur = first_row;

    for (int off = 0; off < a.sub_nnz; off += G) {   // uniform trip count across the warp
        const long long idx = s_begin + off + l;
        int r = -1, c = 0;
        Shfl v = 0;
        if (idx < s_end) {
            r = ld_stream(a.rowind + idx);
            c = ld_stream(a.colind + idx);
            v = ld_stream(a.val + idx);
        }
#pragma unroll
        for (int j0 = 0; j0 < G; j0 += UNROLL) {
            Pack<T, E> b[UNROLL];
            int rr[UNROLL];
            Shfl vv[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                rr[u] = __shfl_sync(FULL, r, j0 + u, G);
                const int cc = __shfl_sync(FULL, c, j0 + u, G);
                vv[u] = __shfl_sync(FULL, v, j0 + u, G);
                if (active && rr[u] >= 0) b[u] = ld_dense<T, E>(Bcol + (long long)cc * a.ldb);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (active && rr[u] >= 0) {
                    if (rr[u] != cur) {
                        coo_flush<T, E>(Ccol + (long long)cur * a.ldc, acc, !(cur == first_row && shared_head),
                                        a.accumulate != 0);
#pragma unroll
                        for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
                        cur = rr[u];
                    }
                    fma_pack<T, E>(acc, b[u], vv[u]);
                }
            }
        }
    }
    if (active && nonempty) {
        coo_flush<T, E>(Ccol + (long long)cur * a.ldc, acc, !(cur == first_row && shared_head) && !shared_tail,
                        a.accumulate != 0);
    }
}

}  // namespace pygim
