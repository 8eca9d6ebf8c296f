// CSR SpMM for sm_100a: C[r, :] = sum_{e in row r} val[e] * B[col[e], :]
//
// Replaces the DPU kernel spmm_default/dpu_kernels/spmm_mul_csr_dpu.c:34-135 (and the grande /
// multigroup copies).  Where a DPU tasklet walks a contiguous row block and issues one MRAM read
// of dense_size*byte_dt bytes per nonzero (:108-126), here:
//
//  * the grid is PERSISTENT (resident warps only); every warp pulls work items from a global,
//    self-resetting (CUDA-graph friendly) ticket counter, so no warp idles while another still owns
//    a long row (the reference's second-level balancing, partition_tsklt_by_nnz_csr,
//    support/partition.c:186-229, made dynamic);
//  * a work item is a group of 1..31 consecutive rows (about 256 nonzeros), or - for rows longer
//    than seg_len - one seg_len-bounded segment of a row; segments come first, longest first;
//  * the warp reads 32*R column indices and values per coalesced evict-first load, D batches ahead
//    of the one being consumed, and hands them round with shuffles: the HBM latency of the index
//    stream is off the critical path of the gathers.  The next ticket and its rowptr entries are
//    fetched before the current item is processed;
//  * a dense row of H elements is covered by G lanes, each moving one 16-byte word (float4,
//    16 x int8, ...), so P = 32/G nonzeros are gathered by every load instruction and UNROLL
//    independent gathers per lane are in flight before the first FMA consumes one.  The gather
//    address is one IMAD.WIDE.U32 (32-bit byte stride, base held in registers);
//  * the P partial sums are combined by an xor-shuffle tree in a fixed order (deterministic), and
//    the row is written once with a streaming store - no host merge (memcpy_2D / memadd_2D,
//    spmm_mul_csr.c:41-86) remains.  With peers set, the row goes to every GPU instead (fused
//    all-gather over NVLink);
//  * a segment publishes its partial sum; the last segment of a row to arrive adds the row's
//    partials in slot order and writes the row (no second kernel, no floating-point atomics);
//  * short-row graphs take csr_stream_rows: the consecutive rows of a ticket are ONE contiguous
//    nonzero stream, read in prefetched batches that ignore row boundaries and walked run by run;
//  * UNIT: when the plan found every stored value equal to one, the value stream is not read.
#pragma once
#include "vec.cuh"

namespace pygim {

constexpr int kMaxPeers = 8;   // GPUs of one NVSwitch box

struct Seg {       // one nnz-bounded piece of a long row
    int long_idx;  // which long row (index into long_rows / long_seg_ptr)
    int start;     // first nonzero (index into colind/val)
    int end;       // one past the last nonzero
    int slot;      // row of the partial buffer this piece writes (slots of one row are consecutive)
};

template <typename T> struct CsrArgs {
    const int *rowptr;
    const int *colind;
    const T *val;
    const T *B;        // dense input, row stride ldb elements
    T *C;              // output, row stride ldc elements
    T *partial;        // [n_seg x ldp] scratch for segment items
    const Seg *segs;
    const int *long_rows;       // [n_long] row id of every long row
    const int *long_seg_ptr;    // [n_long + 1] slots of long row i are [ptr[i], ptr[i+1])
    int *seg_count;             // [col_chunks x n_long] arrival counters, zero between launches
    int n_long;
    unsigned long long *ticket;      // ticket[0]: work counter, ticket[1]: warps that have left the kernel;
                                     // both are zero between launches (the last warp out resets them)
    unsigned int n_warps;            // warps of this launch
    int n_seg;
    int nrows;
    int seg_len;       // rows with more nonzeros than this are handled through segs
    int rows_per_ticket;   // short-row graphs: a ticket covers this many consecutive rows (1..31)
    int n_row_tickets;     // ceil(nrows / rows_per_ticket)
    int nvec;          // words (of E elements) per dense row
    int col_chunks;    // ceil(nvec / G): every item is processed once per chunk of G words
    long long ldb, ldc, ldp;
    unsigned ldb_bytes; // ldb * sizeof(T) (< 4 GiB): 32-bit so a gather address is one IMAD.WIDE.U32
    int accumulate;    // 0: C = A*B, 1: C += A*B
    // fused all-gather (row-sharded multi-GPU): when n_peers > 0 every output row is stored to the same
    // offset of every peer's C (NVLink-mapped pointers, the local one included) instead of a.C; when mc is
    // set it is an NVSwitch multicast mapping of those buffers and ONE multimem.st reaches every GPU.
    T *peers[kMaxPeers];
    T *mc;
    int n_peers;
};

constexpr int kCsrThreads = 256;

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

// store one word of an output row to every destination of the fused all-gather
template <typename T, int E>
__device__ __forceinline__ void st_peers(T *const *peers, int n_peers, T *mc, long long off, const Pack<T, E> &v) {
    if constexpr (sizeof(T) * E == 16) {
        if (mc != nullptr) {
            union { Pack<T, E> p; float4 f; } u;
            u.p = v;
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + off),
                         "f"(u.f.x), "f"(u.f.y), "f"(u.f.z), "f"(u.f.w)
                         : "memory");
            return;
        }
    }
    for (int p = 0; p < n_peers; ++p) st_plain<T, E>(peers[p] + off, v);
}

struct CsrItem {
    int long_idx;      // segment: its long row; rows: -1
    int first;         // segment: slot of the partial buffer; rows: first row of the ticket
    int count;         // rows covered (1 for a segment)
    int rp;            // rows: lane l holds rowptr[first + l] (l <= count); segment: lane 0 start, lane 1 end
    int chunk;         // column chunk
    bool to_partial;
};

// tickets: col_chunks x (n_seg segment items, then n_row_tickets row groups)
template <typename T>
__device__ __forceinline__ CsrItem csr_load_item(const CsrArgs<T> &a, unsigned long long it) {
    CsrItem r;
    const int lane = threadIdx.x & 31;
    const unsigned long long items = (unsigned long long)a.n_seg + (unsigned long long)a.n_row_tickets;
    r.chunk = (int)(it / items);
    const long long k = (long long)(it % items);
    if (k < a.n_seg) {
        const Seg sg = a.segs[k];
        r.long_idx = sg.long_idx;
        r.first = sg.slot;
        r.count = 1;
        r.rp = lane == 0 ? sg.start : sg.end;
        r.to_partial = true;
    } else {
        r.long_idx = -1;
        r.first = (int)(k - a.n_seg) * a.rows_per_ticket;
        r.count = min(a.rows_per_ticket, a.nrows - r.first);
        r.rp = a.rowptr[min(r.first + lane, a.nrows)];
        r.to_partial = false;
    }
    return r;
}

// A segment of a long row has just published its partial sum.  The LAST segment of the row to arrive adds the
// row's partials in slot order (fixed order => bitwise reproducible, no floating-point atomics) and writes the
// final row.  Out of line on purpose: it runs once per segment and must not cost the gather loop registers.
template <typename T, int E, int G>
__device__ __noinline__ void csr_finish_long_row(const CsrArgs<T> &a, int chunk, int long_idx) {
    using Acc = typename Arith<T>::Acc;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int vec = chunk * G + (lane % G);
    const bool active = vec < a.nvec;
    __threadfence();
    __syncwarp();
    int *counter = a.seg_count + (long long)chunk * a.n_long + long_idx;
    int arrived = 0;
    if (lane == 0) {
        __threadfence();          // release: the warp's partial-sum stores (ordered by the barrier above) before the count
        arrived = atomicAdd(counter, 1);
    }
    arrived = __shfl_sync(FULL, arrived, 0);
    const int s0 = a.long_seg_ptr[long_idx], s1 = a.long_seg_ptr[long_idx + 1];
    if (arrived != s1 - s0 - 1) return;
    __threadfence();
    if (lane == 0) *counter = 0;          // ready for the next launch
    if (!(sub == 0 && active)) return;
    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    for (int s = s0; s < s1; ++s)
        add_old<T, E>(acc, ld_cg<T, E>(a.partial + (long long)s * a.ldp + (long long)vec * E));
    const long long off = (long long)a.long_rows[long_idx] * a.ldc + (long long)vec * E;
    if (a.n_peers > 0) {
        st_peers<T, E>(a.peers, a.n_peers, a.mc, off, narrow<T, E>(acc));
    } else {
        if (a.accumulate) add_old<T, E>(acc, ld_plain<T, E>(a.C + off));
        st_stream<T, E>(a.C + off, narrow<T, E>(acc));
    }
}

// R = index entries held per lane per batch (a batch is 32*R nonzeros), D = batches prefetched ahead.
// UNIT: the plan found every stored value equal to one (the value-less adjacency ToSparseTensor yields,
// spmm.py:36-37) - the value stream is then neither loaded nor shuffled and the FMA degenerates to an add;
// results are bit-identical to the general path (x * 1 is exact).
template <typename T, int E, int G, int UNROLL, int R, int D, bool UNIT>
__device__ __forceinline__ void csr_process_range(const CsrArgs<T> &a, int range_start, int range_end, int chunk,
                                                  int dst_row, int long_idx) {
    const bool to_partial = long_idx >= 0;
    using Acc = typename Arith<T>::Acc;
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    constexpr int BATCH = 32 * R;
    constexpr int STEPS = G * R;                 // gather steps per full batch (P nonzeros each)
    constexpr int U = (UNROLL < STEPS) ? UNROLL : STEPS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int vec = chunk * G + (lane % G);
    const bool active = vec < a.nvec;
    const T *Bcol = a.B + (long long)vec * E;
    asm volatile("" : "+l"(Bcol));     // keep the base in a register pair: gather address = one IMAD.WIDE.U32
    const int end = range_end;

    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;

    // index/value stream, prefetched D batches ahead of the gathers (evict-first: read once)
    int nc[D][R];
    Shfl nv[D][R];
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = range_start + d * BATCH + r * 32 + lane;
            nc[d][r] = 0;
            nv[d][r] = 0;
            if (i < end) {
                nc[d][r] = ld_stream(a.colind + i);
                if constexpr (!UNIT) nv[d][r] = ld_stream(a.val + i);
            }
        }
    }
    for (int base = range_start; base < end; base += BATCH) {
        int c[R];
        Shfl v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { c[r] = nc[0][r]; v[r] = nv[0][r]; }
#pragma unroll
        for (int d = 0; d + 1 < D; ++d) {
#pragma unroll
            for (int r = 0; r < R; ++r) { nc[d][r] = nc[d + 1][r]; nv[d][r] = nv[d + 1][r]; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = base + D * BATCH + r * 32 + lane;
            nc[D - 1][r] = 0;
            nv[D - 1][r] = 0;
            if (i < end) {
                nc[D - 1][r] = ld_stream(a.colind + i);
                if constexpr (!UNIT) nv[D - 1][r] = ld_stream(a.val + i);
            }
        }
        const int rem = end - base;
        if (rem >= BATCH) {
            // full batch: STEPS steps of P nonzeros, U gathers in flight per lane.
            // step s covers batch entries s*P .. s*P+P-1 = register s/G, source lane (s%G)*P + sub
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
            // tail batch: per-lane predicate so padded slots never touch B (0 * inf would poison a row)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int left = rem - r * 32;          // entries of register r that are real
                if (left > 0) {
                    const int steps = (min(left, 32) + P - 1) / P;
                    for (int s = 0; s < steps; ++s) {
                        const int src = s * P + sub;
                        const int cc = __shfl_sync(FULL, c[r], src);
                        Shfl vv = (Shfl)1;
                        if constexpr (!UNIT) vv = __shfl_sync(FULL, v[r], src);
                        if (active && src < left) {
                            Pack<T, E> b = ld_dense<T, E>(row_ptr<T>(Bcol, cc, a.ldb_bytes));
                            fma_pack<T, E>(acc, b, vv);
                        }
                    }
                }
            }
        }
    }

    // combine the P interleaved partial sums; fixed tree => bitwise reproducible
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < E; ++k) acc[k] += __shfl_xor_sync(FULL, acc[k], off);
    }
    if (to_partial) {
        if (sub == 0 && active)
            st_plain<T, E>(a.partial + (long long)dst_row * a.ldp + (long long)vec * E, narrow<T, E>(acc));
        csr_finish_long_row<T, E, G>(a, chunk, long_idx);     // rare path, kept out of line
        return;
    }
    if (sub == 0 && active) {
        if (a.n_peers > 0) {
            st_peers<T, E>(a.peers, a.n_peers, a.mc, (long long)dst_row * a.ldc + (long long)vec * E,
                           narrow<T, E>(acc));
        } else {
            T *p = a.C + (long long)dst_row * a.ldc + (long long)vec * E;
            if (a.accumulate) add_old<T, E>(acc, ld_plain<T, E>(p));
            st_stream<T, E>(p, narrow<T, E>(acc));
        }
    }
}

// Write one finished row of a row ticket: combine the P interleaved partial sums (fixed xor tree) and store.
template <typename T, int E, int G>
__device__ __forceinline__ void csr_store_row(const CsrArgs<T> &a, typename Arith<T>::Acc (&acc)[E], int row, int vec,
                                              bool writer) {
    using Acc = typename Arith<T>::Acc;
    constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
        for (int k = 0; k < E; ++k) acc[k] += __shfl_xor_sync(FULL, acc[k], off);
    }
    if (writer) {
        const long long o = (long long)row * a.ldc + (long long)vec * E;
        if (a.n_peers > 0) {
            st_peers<T, E>(a.peers, a.n_peers, a.mc, o, narrow<T, E>(acc));
        } else {
            if (a.accumulate) add_old<T, E>(acc, ld_plain<T, E>(a.C + o));
            st_stream<T, E>(a.C + o, narrow<T, E>(acc));
        }
    }
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
}

// SHORT-ROW graphs: the rows [ja, jb) of a ticket are consecutive, so their nonzeros are ONE contiguous stream.
// It is read in prefetched 32-entry batches that ignore row boundaries (the index-load latency of a row is hidden
// behind the rows before it) and walked run by run - a run being the part of a row inside the batch; row ends
// come from the ticket's rowptr entries held in the lanes (`rp`: lane l holds rowptr[first + l]).
template <typename T, int E, int G, int UT, int D, bool UNIT>
__device__ __forceinline__ void csr_stream_rows(const CsrArgs<T> &a, int first, int ja, int jb, int rp, int chunk) {
    using Acc = typename Arith<T>::Acc;
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int vec = chunk * G + (lane % G);
    const bool active = vec < a.nvec;
    const bool writer = active && sub == 0;
    const T *Bcol = a.B + (long long)vec * E;
    asm volatile("" : "+l"(Bcol));
    const int s0 = __shfl_sync(FULL, rp, ja), s1 = __shfl_sync(FULL, rp, jb);

    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    int cur = ja;                                        // row (relative to `first`) being accumulated
    int boundary = __shfl_sync(FULL, rp, cur + 1);       // one past its last nonzero

    int nc[D];
    Shfl nv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int i = s0 + d * 32 + lane;
        nc[d] = 0;
        nv[d] = 0;
        if (i < s1) {
            nc[d] = ld_stream(a.colind + i);
            if constexpr (!UNIT) nv[d] = ld_stream(a.val + i);
        }
    }
    for (int base = s0; base < s1; base += 32) {
        const int c = nc[0];
        const Shfl v = nv[0];
#pragma unroll
        for (int d = 0; d + 1 < D; ++d) { nc[d] = nc[d + 1]; nv[d] = nv[d + 1]; }
        {
            const int i = base + D * 32 + lane;
            nc[D - 1] = 0;
            nv[D - 1] = 0;
            if (i < s1) {
                nc[D - 1] = ld_stream(a.colind + i);
                if constexpr (!UNIT) nv[D - 1] = ld_stream(a.val + i);
            }
        }
        const int left = min(32, s1 - base);
        int pos = 0;
        while (pos < left) {                             // warp-uniform
            while (boundary <= base + pos) {             // rows that ended (or are empty): write them out
                csr_store_row<T, E, G>(a, acc, first + cur, vec, writer);
                ++cur;
                boundary = __shfl_sync(FULL, rp, cur + 1);
            }
            const int run_end = min(left, boundary - base);
#pragma unroll 1
            for (int s = pos; s < run_end; s += P * UT) {
                Pack<T, E> b[UT];
#pragma unroll
                for (int u = 0; u < UT; ++u) {
                    const int src = s + u * P + sub;
                    const int cc = __shfl_sync(FULL, c, src & 31);
                    if (active && src < run_end) b[u] = ld_dense<T, E>(row_ptr<T>(Bcol, cc, a.ldb_bytes));
                }
#pragma unroll
                for (int u = 0; u < UT; ++u) {
                    const int src = s + u * P + sub;
                    Shfl vv = (Shfl)1;
                    if constexpr (!UNIT) vv = __shfl_sync(FULL, v, src & 31);
                    if (active && src < run_end) fma_pack<T, E>(acc, b[u], vv);
                }
            }
            pos = run_end;
        }
    }
    while (cur < jb) {                                   // the last row with data and any trailing empty rows
        csr_store_row<T, E, G>(a, acc, first + cur, vec, writer);
        ++cur;
    }
}

// Persistent grid: gridDim.x = resident blocks of the device.  Tickets run over
// col_chunks * (n_seg + n_row_tickets) items, column chunk outermost.
template <typename T, int E, int G, int UNROLL, int MIN_BLOCKS, int R, int D, bool UNIT, bool STREAM = false>
__global__ void __launch_bounds__(kCsrThreads, MIN_BLOCKS) csr_spmm_kernel(const __grid_constant__ CsrArgs<T> a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned long long total =
        (unsigned long long)a.col_chunks * ((unsigned long long)a.n_seg + (unsigned long long)a.n_row_tickets);

    auto take_ticket = [&]() -> unsigned long long {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1ULL);
        return __shfl_sync(FULL, t, 0);
    };

    unsigned long long it = take_ticket();
    CsrItem cur;
    if (it < total) cur = csr_load_item<T>(a, it);
    while (it < total) {
        // look ahead: the next ticket and its row bounds are in flight while this item is processed
        const unsigned long long nit = take_ticket();
        CsrItem nxt;
        if (nit < total) nxt = csr_load_item<T>(a, nit);
        if (STREAM && !cur.to_partial) {
            // rows longer than seg_len are covered by their segments: stream the row blocks between them
            const int deg = __shfl_down_sync(FULL, cur.rp, 1) - cur.rp;
            unsigned long_rows = __ballot_sync(FULL, lane < cur.count && deg > a.seg_len);
            int ja = 0;
            while (ja < cur.count) {
                const unsigned rest = long_rows >> ja;
                const int jb = rest ? ja + (__ffs(rest) - 1) : cur.count;
                if (jb > ja)
                    csr_stream_rows<T, E, G, (UNROLL < 4 ? UNROLL : 4), (D < 2 ? 2 : D), UNIT>(a, cur.first, ja, jb, cur.rp,
                                                                                             cur.chunk);
                ja = jb + 1;
            }
        } else {
            for (int j = 0; j < cur.count; ++j) {
                const int start = __shfl_sync(FULL, cur.rp, j);
                const int end = __shfl_sync(FULL, cur.rp, j + 1);
                // rows longer than seg_len are covered by their segments + the last-arriver merge
                if (cur.to_partial || end - start <= a.seg_len)
                    csr_process_range<T, E, G, UNROLL, R, D, UNIT>(a, start, end, cur.chunk, cur.first + j, cur.long_idx);
            }
        }
        it = nit;
        cur = nxt;
    }
    release_tickets(a.ticket, a.n_warps);
}

}  // namespace pygim
