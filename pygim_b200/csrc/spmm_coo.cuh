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
//  * the grid is persistent; warps draw chunks of chunk_nnz consecutive nonzeros from a ticket
//    counter (same scheme as the CSR kernel);
//  * a warp streams its chunk 32*R triples (row, col, val) at a time - coalesced evict-first
//    loads, prefetched D batches ahead.  If the whole batch belongs to the row being accumulated
//    (the common case on dense-ish graphs) it is gathered exactly like a CSR batch: G lanes x 16
//    bytes per dense row, P = 32/G nonzeros per load instruction, UNROLL gathers in flight;
//  * otherwise the batch is walked run by run: row runs are found with a ballot over the row
//    stream, and at every row change the P interleaved partial sums are combined by the same
//    fixed xor-shuffle tree and flushed - a segmented reduction in registers and shuffles;
//  * rows that lie wholly inside a chunk are written with plain stores; only a chunk's first /
//    last row, and only if the neighbouring nonzero really has the same row, is combined with
//    atomics (integer atomics are exact; float atomics commute up to rounding - see DESIGN.md).
//
// C is zero-filled by the launcher first (the reference's torch::zeros, pytorch_api.cpp:357-358)
// unless accumulate is set.
#pragma once
#include "vec.cuh"

namespace pygim {

// Every warp calls this once, after it drew its last (failing) ticket: the last warp out zeroes the counters, so
// the plan needs no host-side bookkeeping between launches and a launch can be replayed from a CUDA graph.
__device__ __forceinline__ void release_tickets(unsigned long long *ticket, unsigned int n_warps) {
    if ((threadIdx.x & 31) == 0) {
        __threadfence();
        const unsigned long long left = atomicAdd(ticket + 1, 1ULL);
        if (left == (unsigned long long)n_warps - 1ULL) {
            ticket[0] = 0ULL;
            ticket[1] = 0ULL;
            __threadfence();
        }
    }
}

template <typename T> struct CooArgs {
    const int *rowind;
    const int *colind;
    const T *val;
    const T *B;
    T *C;
    unsigned long long *ticket;      // same self-resetting counters as the CSR kernel (spmm_csr.cuh)
    unsigned int n_warps;
    long long nnz;
    long long n_chunks;
    long long ldb, ldc;
    unsigned ldb_bytes; // ldb * sizeof(T) (< 4 GiB)
    int chunk_nnz;     // nonzeros per work item (multiple of 32)
    int nvec;          // words (of E elements) per dense row
    int col_chunks;    // ceil(nvec / G)
    int accumulate;
    int all_atomic;    // the stream is not row-major sorted: no row is owned by one warp, every flush is an atomic add
};

constexpr int kCooThreads = 256;

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

// Combine the P interleaved partial sums (all lanes take part) and write / add the row.
template <typename T, int E, int G>
__device__ __forceinline__ void coo_flush_row(T *Crow, typename Arith<T>::Acc (&acc)[E], bool writer, bool exclusive,
                                              bool accumulate) {
    using Acc = typename Arith<T>::Acc;
    constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < E; ++k) acc[k] += __shfl_xor_sync(FULL, acc[k], off);
    }
    if (writer) {
        if (exclusive) {
            if (accumulate) add_old<T, E>(acc, ld_plain<T, E>(Crow));
            st_plain<T, E>(Crow, narrow<T, E>(acc));
        } else {
#pragma unroll
            for (int k = 0; k < E; ++k) atomic_add_elem<T>(Crow + k, acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
}

template <typename T, int E, int G, int UNROLL, int R, int D, bool UNIT>
__device__ __forceinline__ void coo_process_chunk(const CooArgs<T> &a, long long cs, long long ce, int chunk) {
    using Acc = typename Arith<T>::Acc;
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    constexpr int BATCH = 32 * R;
    constexpr int STEPS = G * R;
    constexpr int U = (UNROLL < STEPS) ? UNROLL : STEPS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int vec = chunk * G + (lane % G);
    const bool active = vec < a.nvec;
    const bool writer = active && sub == 0;
    const bool accumulate = a.accumulate != 0;
    const T *Bcol = a.B + (long long)vec * E;
    asm volatile("" : "+l"(Bcol));     // keep the base in a register pair: gather address = one IMAD.WIDE.U32
    T *Ccol = a.C + (long long)vec * E;

    // does the neighbouring nonzero belong to the same row as our first / last one?
    const int first_row = a.rowind[cs];
    const bool head_shared = a.all_atomic || (cs > 0 && a.rowind[cs - 1] == first_row);
    const bool tail_shared = a.all_atomic || (ce < a.nnz && a.rowind[ce] == a.rowind[ce - 1]);

    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    int cur = first_row;

    int nr[D][R], nc[D][R];
    Shfl nv[D][R];
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long i = cs + d * BATCH + r * 32 + lane;
            nr[d][r] = -1; nc[d][r] = 0; nv[d][r] = 0;
            if (i < ce) {
                nr[d][r] = ld_stream(a.rowind + i);
                nc[d][r] = ld_stream(a.colind + i);
                if constexpr (!UNIT) nv[d][r] = ld_stream(a.val + i);
            }
        }
    }
    for (long long base = cs; base < ce; base += BATCH) {
        int rr[R], c[R];
        Shfl v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { rr[r] = nr[0][r]; c[r] = nc[0][r]; v[r] = nv[0][r]; }
#pragma unroll
        for (int d = 0; d + 1 < D; ++d) {
#pragma unroll
            for (int r = 0; r < R; ++r) { nr[d][r] = nr[d + 1][r]; nc[d][r] = nc[d + 1][r]; nv[d][r] = nv[d + 1][r]; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long i = base + (long long)D * BATCH + r * 32 + lane;
            nr[D - 1][r] = -1; nc[D - 1][r] = 0; nv[D - 1][r] = 0;
            if (i < ce) {
                nr[D - 1][r] = ld_stream(a.rowind + i);
                nc[D - 1][r] = ld_stream(a.colind + i);
                if constexpr (!UNIT) nv[D - 1][r] = ld_stream(a.val + i);
            }
        }
        const long long rem = ce - base;
        // whole batch inside the current row?  (sorted stream: first and last entry decide)
        const int r_first = __shfl_sync(FULL, rr[0], 0);
        const int r_last = __shfl_sync(FULL, rr[R - 1], 31);
        if (rem >= BATCH && r_first == cur && r_last == cur && !a.all_atomic) {
#pragma unroll
            for (int s0 = 0; s0 < STEPS; s0 += U) {
                Pack<T, E> b[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int cc = __shfl_sync(FULL, c[(s0 + u) / G], ((s0 + u) % G) * P + sub);
                    if (active) b[u] = ld_dense<T, E>(row_ptr<T>(Bcol, cc, a.ldb_bytes));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    Shfl vv = (Shfl)1;
                    if constexpr (!UNIT) vv = __shfl_sync(FULL, v[(s0 + u) / G], ((s0 + u) % G) * P + sub);
                    if (active) fma_pack<T, E>(acc, b[u], vv);
                }
            }
        } else {
            // walk the batch run by run (a run = consecutive entries of one row)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long long left64 = rem - r * 32;
                const int left = left64 > 32 ? 32 : (int)left64;       // real entries of register r
                int pos = 0;
                while (pos < left) {                                   // warp-uniform
                    const int row_here = __shfl_sync(FULL, rr[r], pos);
                    if (row_here != cur) {
                        coo_flush_row<T, E, G>(Ccol + (long long)cur * a.ldc, acc, writer,
                                               !(cur == first_row && head_shared) && !a.all_atomic, accumulate);
                        cur = row_here;
                    }
                    const unsigned differs = __ballot_sync(FULL, lane >= pos && lane < left && rr[r] != row_here);
                    const int run_end = differs ? (__ffs(differs) - 1) : left;
                    for (int s = pos; s < run_end; s += P) {
                        const int src = s + sub;
                        const int cc = __shfl_sync(FULL, c[r], src & 31);
                        Shfl vv = (Shfl)1;
                        if constexpr (!UNIT) vv = __shfl_sync(FULL, v[r], src & 31);
                        if (active && src < run_end) {
                            Pack<T, E> b = ld_dense<T, E>(row_ptr<T>(Bcol, cc, a.ldb_bytes));
                            fma_pack<T, E>(acc, b, vv);
                        }
                    }
                    pos = run_end;
                }
            }
        }
    }
    coo_flush_row<T, E, G>(Ccol + (long long)cur * a.ldc, acc, writer,
                           !(cur == first_row && head_shared) && !tail_shared, accumulate);
}

// Persistent grid; tickets run over col_chunks * n_chunks items, column chunk outermost.
template <typename T, int E, int G, int UNROLL, int MIN_BLOCKS, int R, int D, bool UNIT>
__global__ void __launch_bounds__(kCooThreads, MIN_BLOCKS) coo_spmm_kernel(const CooArgs<T> a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned long long total = (unsigned long long)a.col_chunks * (unsigned long long)a.n_chunks;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1ULL);
        t = __shfl_sync(FULL, t, 0);
        if (t >= total) break;
        const int chunk = (int)(t / (unsigned long long)a.n_chunks);
        const long long k = (long long)(t % (unsigned long long)a.n_chunks);
        const long long cs = k * a.chunk_nnz;
        long long ce = cs + a.chunk_nnz;
        if (ce > a.nnz) ce = a.nnz;
        coo_process_chunk<T, E, G, UNROLL, R, D, UNIT>(a, cs, ce, chunk);
    }
    release_tickets(a.ticket, a.n_warps);
}

}  // namespace pygim
