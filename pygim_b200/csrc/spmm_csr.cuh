// CSR SpMM for sm_100a: C[r, :] = sum_{e in row r} val[e] * B[col[e], :]
//
// Replaces the DPU kernel spmm_default/dpu_kernels/spmm_mul_csr_dpu.c:34-135 (and the grande /
// multigroup copies).  Where a DPU tasklet walks a contiguous row block and issues one MRAM read
// of dense_size*byte_dt bytes per nonzero (:108-126), here:
//
//  * one warp owns one work item: a whole row, or - for rows longer than seg_len - one
//    seg_len-bounded segment of a row (the nnz-balanced second level of the reference's
//    partition_tsklt_by_nnz_csr, support/partition.c:186-229, taken to exact nnz granularity);
//  * the warp reads 32 column indices and 32 values with one coalesced, evict-first load each and
//    hands them round with shuffles;
//  * a dense row of H elements is covered by G lanes, each moving one 16-byte word (float4,
//    16 x int8, ...), so P = 32/G nonzeros are gathered by every load instruction and up to UNROLL
//    independent gathers per lane are in flight before the first FMA consumes one;
//  * the P partial sums are combined by an xor-shuffle tree in a fixed order (deterministic), and
//    the row is written once with a streaming store - no host merge (memcpy_2D / memadd_2D,
//    spmm_mul_csr.c:41-86) remains.
//
// Segment items write to a partial buffer; csr_fixup_kernel adds a row's partials in segment order.
#pragma once
#include "vec.cuh"

namespace pygim {

struct Seg {       // one nnz-bounded piece of a long row
    int row;
    int start;     // first nonzero (index into colind/val)
    int end;       // one past the last nonzero
    int slot;      // row of the partial buffer this piece writes
};

template <typename T> struct CsrArgs {
    const int *rowptr;
    const int *colind;
    const T *val;
    const T *B;        // dense input, row stride ldb elements
    T *C;              // output, row stride ldc elements
    T *partial;        // [n_seg x ldp] scratch for segment items
    const Seg *segs;
    int n_seg;
    int nrows;
    int seg_len;       // rows with more nonzeros than this are handled through segs
    int nvec;          // words (of E elements) per dense row
    long long ldb, ldc, ldp;
    int accumulate;    // 0: C = A*B, 1: C += A*B
};

constexpr int kCsrWarpsPerBlock = 8;

template <typename T, int E, int G, int UNROLL>
__device__ __forceinline__ void csr_gather_range(const CsrArgs<T> &a, int start, int end, const T *Bcol, bool active,
                                                 typename Arith<T>::Acc (&acc)[E]) {
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;

    for (int base = start; base < end; base += 32) {
        const int idx = base + lane;
        int c = 0;
        Shfl v = 0;
        if (idx < end) {
            c = ld_stream(a.colind + idx);
            v = ld_stream(a.val + idx);
        }
        const int rem = end - base;
        if (rem >= 32) {
            // full batch: G steps of P nonzeros, UNROLL gathers in flight per lane
#pragma unroll
            for (int s0 = 0; s0 < G; s0 += UNROLL) {
                Pack<T, E> b[UNROLL];
                Shfl vv[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int src = (s0 + u) * P + sub;
                    const int cc = __shfl_sync(FULL, c, src);
                    vv[u] = __shfl_sync(FULL, v, src);
                    if (active) b[u] = ld_dense<T, E>(Bcol + (long long)cc * a.ldb);
                }
                if (active) {
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) fma_pack<T, E>(acc, b[u], vv[u]);
                }
            }
        } else {
            // tail batch: per-lane predicate so padded slots never touch B (0 * inf would poison a row)
            const int steps = (rem + P - 1) / P;
            for (int s = 0; s < steps; ++s) {
                const int src = s * P + sub;
                const int cc = __shfl_sync(FULL, c, src);
                const Shfl vv = __shfl_sync(FULL, v, src);
                if (active && src < rem) {
                    Pack<T, E> b = ld_dense<T, E>(Bcol + (long long)cc * a.ldb);
                    fma_pack<T, E>(acc, b, vv);
                }
            }
        }
    }
}

// grid.x = ceil((n_seg + nrows) / kCsrWarpsPerBlock); grid.y = column chunks of G words (only > 1 when G == 32)
template <typename T, int E, int G>
__global__ void __launch_bounds__(kCsrWarpsPerBlock * 32) csr_spmm_kernel(const CsrArgs<T> a) {
    using Acc = typename Arith<T>::Acc;
    constexpr int P = 32 / G;
    constexpr int UNROLL = (G < 8) ? G : 8;
    constexpr unsigned FULL = 0xffffffffu;

    const long long item = (long long)blockIdx.x * kCsrWarpsPerBlock + (threadIdx.x >> 5);
    if (item >= (long long)a.n_seg + a.nrows) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int vec = blockIdx.y * G + (lane % G);
    const bool active = vec < a.nvec;

    int start, end;
    T *dst;
    bool accumulate = a.accumulate != 0;
    if (item < a.n_seg) {
        // long-row segments come first so the biggest items are scheduled earliest
        const Seg sg = a.segs[item];
        start = sg.start;
        end = sg.end;
        dst = a.partial + (long long)sg.slot * a.ldp;
        accumulate = false;
    } else {
        const int row = (int)(item - a.n_seg);
        start = a.rowptr[row];
        end = a.rowptr[row + 1];
        if (end - start > a.seg_len) return;   // handled by its segments + fix-up
        dst = a.C + (long long)row * a.ldc;
    }

    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;

    csr_gather_range<T, E, G, UNROLL>(a, start, end, a.B + (long long)vec * E, active, acc);

    // combine the P interleaved partial sums; fixed tree => bitwise reproducible
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < E; ++k) acc[k] += __shfl_xor_sync(FULL, acc[k], off);
    }
    if (sub == 0 && active) {
        T *p = dst + (long long)vec * E;
        if (accumulate) add_old<T, E>(acc, ld_plain<T, E>(p));
        st_stream<T, E>(p, narrow<T, E>(acc));
    }
}

// One block per long row: C[row, :] (+)= sum over the row's segments, in segment order.
template <typename T> struct FixupArgs {
    const T *partial;
    T *C;
    const int *long_rows;      // [n_long] row ids
    const int *long_seg_ptr;   // [n_long + 1] slots of row i are [ptr[i], ptr[i+1])
    long long ldp, ldc;
    int ncols;
    int accumulate;
};

template <typename T> __global__ void csr_fixup_kernel(const FixupArgs<T> a) {
    using Acc = typename Arith<T>::Acc;
    const int i = blockIdx.x;
    const int row = a.long_rows[i];
    const int s0 = a.long_seg_ptr[i], s1 = a.long_seg_ptr[i + 1];
    for (int c = threadIdx.x; c < a.ncols; c += blockDim.x) {
        Acc acc[1] = {(Acc)0};
        for (int s = s0; s < s1; ++s) {
            Pack<T, 1> p;
            p.e[0] = a.partial[(long long)s * a.ldp + c];
            add_old<T, 1>(acc, p);
        }
        T *dst = a.C + (long long)row * a.ldc + c;
        if (a.accumulate) {
            Pack<T, 1> o;
            o.e[0] = *dst;
            add_old<T, 1>(acc, o);
        }
        *dst = (T)acc[0];
    }
}

}  // namespace pygim
