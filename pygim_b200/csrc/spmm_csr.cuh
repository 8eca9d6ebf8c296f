// CSR SpMM for sm_100a: C[r, :] = sum_{e in row r} val[e] * B[col[e], :]
//
// Replaces the DPU kernel spmm_default/dpu_kernels/spmm_mul_csr_dpu.c:34-135 (and the grande /
// multigroup copies).  Where a DPU tasklet walks a contiguous row block and issues one MRAM read
// of dense_size*byte_dt bytes per nonzero (:108-126), here:
//
//  * WORK ITEMS are built once per plan, in row order: a group of 1..31 consecutive rows holding
//    about item_nnz nonzeros, or - for rows longer than seg_len - one seg_len-bounded segment of a
//    row (the reference's second-level balancing, partition_tsklt_by_nnz_csr,
//    support/partition.c:186-229).  Consecutive items are bundled into SUPERTICKETS of near-equal
//    nnz;
//  * the grid is PERSISTENT.  Default plans are ONE queue of items (exactly balanced; the first round is dealt
//    statically, warp w takes item w).  Plans whose rows were reordered for locality (pygim_b200/reorder.py) and
//    whose dense rows are >= 256 bytes are SM-AFFINE: SM s owns supertickets s, s + #SMs, s + 2 #SMs ... and every
//    warp resident on it drains the same superticket (one atomic per item on the superticket's own counter), so
//    the warps that share an SM's L1 work on neighbouring rows and the dense rows they gather are L1 hits instead
//    of L2 -> SM traffic.  Warps that run out of home work STEAL items from any unfinished superticket, so no
//    warp idles while another still owns a long row;
//  * citation-shaped graphs (mean degree < 12) are split at plan time: rows of <= 8 nonzeros go to
//    csr_tiny_rows_kernel (one lane group per row, a grid as wide as the matrix), every other row is a list of
//    <= seg_len pieces drawn longest first by this kernel (STREAM == 2);
//  * a dense row of H elements is covered by G lanes, each moving one 16-byte word (float4,
//    16 x int8, ...), so P = 32/G nonzeros are gathered by every load instruction.  Every lane
//    group reads the column indices (and values) of ITS OWN four consecutive nonzeros as one
//    16-byte vector load, one round ahead of the gathers that consume them - no shuffles on the
//    gather path, 4 or 8 independent gathers per lane in flight.  The gather address is one
//    IMAD.WIDE.U32 (32-bit byte stride, base held in registers);
//  * the P partial sums are combined by an xor-shuffle tree in a fixed order (deterministic), and
//    the row is written once with a streaming store - no host merge (memcpy_2D / memadd_2D,
//    spmm_mul_csr.c:41-86) remains.  The store is also where the rest of a conv layer is fused:
//    row un-permutation (row_map), integer -> float de-quantisation (scale), the (1+eps)*x
//    residual, and - with peers set - the all-gather (the row goes to every GPU over NVLink);
//  * a segment publishes its partial sum with a RELEASE (no L1 invalidation); the last segment of
//    a row to arrive adds the row's partials in slot order and writes the row (no second kernel,
//    no floating-point atomics);
//  * short-row graphs take csr_stream_rows: the consecutive rows of an item are ONE contiguous
//    nonzero stream, read in prefetched batches that ignore row boundaries and walked run by run;
//  * UNIT: when the plan found every stored value equal to one, the value stream is not read.
#pragma once
#include "vec.cuh"

namespace pygim {

constexpr int kMaxPeers = 8;   // GPUs of one NVSwitch box

struct Seg {       // one nnz-bounded piece of a long row
    int long_idx;  // which long row (index into long_rows / long_seg_ptr); < 0: the piece is the WHOLE row ~long_idx
    int start;     // first nonzero (index into colind/val)
    int end;       // one past the last nonzero
    int slot;      // row of the partial buffer this piece writes (slots of one row are consecutive)
};

static_assert(sizeof(Seg) == 16, "segment descriptors are read as one 16-byte word");

// what the row store does besides writing the sum (all optional, all decided per launch)
struct Epilogue {
    const int *row_map;        // plan row r is row row_map[r] of the result (row reordering); null = identity
    const float *scale;        // device scalar: the result is FLOAT32, (float)sum * scale[0] (symmetric_dequantize)
    const float *residual;     // float32 [rows x H], row stride ld_res: result += coeff * residual (GIN's (1+eps) x)
    long long ld_res;
    float coeff;
    // fused all-gather (row-sharded multi-GPU): when n_peers > 0 every output row is stored to the same offset of
    // every peer's C (NVLink-mapped pointers, the local one included) instead of C; when mc is set it is an
    // NVSwitch multicast mapping of those buffers and ONE multimem.st reaches every GPU.  peer_mask (optional,
    // one byte per plan row) restricts a row to the peers whose bit is set (halo exchange).
    void *peers[kMaxPeers];
    void *mc;
    const unsigned char *peer_mask;
    int n_peers;
    // arrival flags: the last warp of the launch stores `epoch` into flags[p][my_rank] of every peer
    int *flags[kMaxPeers];
    int my_rank;
    int epoch;
};

template <typename T> struct CsrArgs {
    const int *rowptr;
    const int *colind;
    const T *val;
    const T *B;        // dense input, row stride ldb elements
    void *C;           // output (T, or float when epi.scale is set), row stride ldc elements
    T *partial;        // [n_seg x ldp] scratch for segment items (ldp covers the whole dense row)
    const Seg *segs;
    const int4 *supers;         // [n_super] supertickets.  x >= 0: rows [x, x + y) dealt z rows per item (w items);
                                //            x < 0: the w segments ~x, ~x + 1, ... of long rows
    const int *long_rows;       // [n_long] row id of every long row
    const int *long_seg_ptr;    // [n_long + 1] slots of long row i are [ptr[i], ptr[i+1])
    int *super_cnt;             // [n_super x col_chunks] items drawn per (superticket, column chunk); zero at rest
    int *seg_count;             // [col_chunks x n_long] arrival counters, zero at rest
    unsigned int *warps_out;    // warps that have left the kernel; zero at rest (the last one resets everything)
                                // warps_out[1]: (superticket, chunk) units whose items have ALL been drawn
    unsigned int n_warps;       // warps of this launch
    int n_super, n_seg, n_long;
    int nrows;
    int seg_len;       // rows with more nonzeros than this are covered by segments (row items skip them)
    int nvec;          // words (of E elements) per dense row
    int col_chunks;    // ceil(nvec / G): every item is processed once per chunk of G words
    long long nnz_total;        // entries of colind/val (bounds of the vector index loads)
    int idx_mis;       // (address of colind / 4) mod 4 when val is misaligned the same way, else 4 = scalar index loads
    long long ldb, ldc, ldp;
    unsigned ldb_bytes; // ldb * sizeof(T) (< 4 GiB): 32-bit so a gather address is one IMAD.WIDE.U32
    int accumulate;    // 0: C = A*B, 1: C += A*B
    Epilogue epi;
};

// ---------------------------------------------------------------------------------------------- row store
__device__ __forceinline__ void st_multimem(void *mc, long long byte_off, const float4 &f) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"((char *)mc + byte_off),
                 "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w)
                 : "memory");
}

template <typename T, int E> struct AccPack { typename Arith<T>::Acc v[E]; };

// FLOAT32 result of one word: de-quantise (scale) and/or add coeff * residual.  One multiply, one multiply, one add,
// never contracted, so the result equals torch's `out_q * scale + coeff * x` bit for bit
// (models/quantize.py:40-42, pyg_gin_conv.py:84-86).  Out of line: runs once per row.
template <typename T, int E>
__device__ __noinline__ void csr_emit_float(const CsrArgs<T> &a, AccPack<T, E> acc, int row, long long orow, int vec) {
    const Epilogue &ep = a.epi;
    const long long off = orow * a.ldc + (long long)vec * E;
    float o[E];
    const float s = ep.scale ? *ep.scale : 1.0f;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const float v = (float)(T)acc.v[k];
        o[k] = ep.scale ? __fmul_rn(v, s) : v;
    }
    if (ep.residual) {
        const float *r = ep.residual + orow * ep.ld_res + (long long)vec * E;
#pragma unroll
        for (int k = 0; k < E; ++k) o[k] = __fadd_rn(o[k], __fmul_rn(ep.coeff, r[k]));
    }
    float *Cf = static_cast<float *>(a.C);
    if (ep.n_peers > 0) {
        const unsigned m = ep.peer_mask ? ep.peer_mask[row] : 0xffu;
        if constexpr (E % 4 == 0) {
            if (ep.mc != nullptr && m == 0xffu) {
#pragma unroll
                for (int q = 0; q < E / 4; ++q)
                    st_multimem(ep.mc, (off + 4 * q) * 4, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
                return;
            }
        }
        for (int p = 0; p < ep.n_peers; ++p) {
            if (!((m >> p) & 1u)) continue;
            float *dst = static_cast<float *>(ep.peers[p]) + off;
#pragma unroll
            for (int k = 0; k < E; ++k) dst[k] = o[k];
        }
        return;
    }
    if constexpr (E % 4 == 0) {
#pragma unroll
        for (int q = 0; q < E / 4; ++q)
            __stcs(reinterpret_cast<float4 *>(Cf + off) + q, make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) __stcs(Cf + off + k, o[k]);
    }
}

// Store the E finished elements of (plan row `row`, word `vec`).  `acc` holds the full sums.
template <typename T, int E>
__device__ __forceinline__ void csr_emit(const CsrArgs<T> &a, typename Arith<T>::Acc (&acc)[E], int row, int vec) {
    const Epilogue &ep = a.epi;
    const long long orow = ep.row_map ? (long long)ep.row_map[row] : (long long)row;
    if (ep.scale != nullptr || ep.residual != nullptr) {
        AccPack<T, E> pk;
#pragma unroll
        for (int k = 0; k < E; ++k) pk.v[k] = acc[k];
        csr_emit_float<T, E>(a, pk, row, orow, vec);
        return;
    }
    const long long off = orow * a.ldc + (long long)vec * E;
    T *C = static_cast<T *>(a.C);
    if (ep.n_peers > 0) {
        const unsigned m = ep.peer_mask ? ep.peer_mask[row] : 0xffu;
        const Pack<T, E> v = narrow<T, E>(acc);
        if constexpr (sizeof(T) * E == 16) {
            if (ep.mc != nullptr && m == 0xffu) {
                union { Pack<T, E> p; float4 f; } u;
                u.p = v;
                st_multimem(ep.mc, off * (long long)sizeof(T), u.f);
                return;
            }
        }
        for (int p = 0; p < ep.n_peers; ++p)
            if ((m >> p) & 1u) st_plain<T, E>(static_cast<T *>(ep.peers[p]) + off, v);
        return;
    }
    if (a.accumulate) add_old<T, E>(acc, ld_plain<T, E>(C + off));
    st_stream<T, E>(C + off, narrow<T, E>(acc));
}

// Combine the P interleaved partial sums (fixed xor tree => bitwise reproducible) and store the row.
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
    if (writer) csr_emit<T, E>(a, acc, row, vec);
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
}

// A segment of a long row has just published its partial sum.  The LAST segment of the row to arrive adds the
// row's partials in slot order (fixed order => bitwise reproducible, no floating-point atomics) and writes the
// final row.  Out of line on purpose: it runs once per segment and must not cost the gather loop registers.
// Publishing is a RELEASE (MEMBAR without CCTL.IVALL): the SM's L1 - the whole point of the SM-affine schedule -
// survives; only the one warp that merges a row pays an ACQUIRE, and it reads the partials past L1 anyway.
template <typename T, int E, int G>
__device__ __noinline__ void csr_finish_long_row(const CsrArgs<T> &a, int chunk, int long_idx) {
    using Acc = typename Arith<T>::Acc;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int vec = chunk * G + (lane % G);
    const bool active = vec < a.nvec;
    __syncwarp();                 // the warp's partial-sum stores are ordered before lane 0's release
    int *counter = a.seg_count + (long long)chunk * a.n_long + long_idx;
    int arrived = 0;
    if (lane == 0)
        asm volatile("atom.add.release.gpu.global.s32 %0, [%1], 1;" : "=r"(arrived) : "l"(counter) : "memory");
    arrived = __shfl_sync(FULL, arrived, 0);
    const int s0 = a.long_seg_ptr[long_idx], s1 = a.long_seg_ptr[long_idx + 1];
    if (arrived != s1 - s0 - 1) return;
    asm volatile("fence.acquire.gpu;" ::: "memory");
    if (lane == 0) *counter = 0;          // ready for the next launch
    if (!(sub == 0 && active)) return;
    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    for (int s = s0; s < s1; ++s) add_old<T, E>(acc, ld_cg<T, E>(a.partial + (long long)s * a.ldp + (long long)vec * E));
    csr_emit<T, E>(a, acc, a.long_rows[long_idx], vec);
}

// ---------------------------------------------------------------------------------------------- gather loops
// one nonzero: gather word `vec` of dense row `col` and accumulate
template <typename T, int E, bool UNIT>
__device__ __forceinline__ void csr_one(const int *colind, const T *val, unsigned ldb_bytes, int i, const T *Bcol,
                                        typename Arith<T>::Acc (&acc)[E]) {
    using Shfl = typename Arith<T>::Shfl;
    const int c = __ldg(colind + i);
    Shfl v = (Shfl)1;
    if constexpr (!UNIT) v = (Shfl)__ldg(val + i);
    const Pack<T, E> b = ld_dense<T, E>(row_ptr<T>(Bcol, c, ldb_bytes));
    fma_pack<T, E>(acc, b, v);
}

// values of the four nonzeros of one aligned word
template <typename T>
__device__ __forceinline__ void ld_val4(const T *p, typename Arith<T>::Shfl (&v)[4]) {
    using Shfl = typename Arith<T>::Shfl;
    if constexpr (sizeof(T) == 8) {
        const Pack<T, 2> v0 = ld_dense<T, 2>(p), v1 = ld_dense<T, 2>(p + 2);
        v[0] = (Shfl)v0.e[0]; v[1] = (Shfl)v0.e[1]; v[2] = (Shfl)v1.e[0]; v[3] = (Shfl)v1.e[1];
    } else {
        const Pack<T, 4> vv = ld_dense<T, 4>(p);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (Shfl)vv.e[k];
    }
}

// acc += sum over the nonzeros [start, end) of val * B[col, word `vec`].
// The range is cut at the 16-byte boundaries of the index array: the (at most six) nonzeros before the first and
// after the last boundary are handled one by one, the words in between are dealt round-robin to the P lane groups,
// each group reading the four column indices (and values) of its word with ONE vector load.  NV words per group
// are in flight: 4*NV independent gathers per lane before the first FMA, and the indices of the following NV words
// are already loading.
template <typename T, int E, int G, int NV, bool UNIT>
__device__ __forceinline__ AccPack<T, E> csr_accumulate(const int *colind, const T *val, const T *Bcol, unsigned ldb_bytes,
                                                        int idx_mis, int start, int end, bool active) {
    using Acc = typename Arith<T>::Acc;
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    const int sub = (threadIdx.x & 31) / G;
    AccPack<T, E> out;
    Acc (&acc)[E] = out.v;
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    if (!active) return out;
    if (idx_mis >= 4) {           // colind / val not co-aligned (foreign views): element by element
        for (int i = start + sub; i < end; i += P) csr_one<T, E, UNIT>(colind, val, ldb_bytes, i, Bcol, acc);
        return out;
    }
    {   // the row's index (and value) stream comes from HBM: pull it into L2 now, 128 bytes per lane and instruction,
        // so the vector loads below - issued only NV words ahead of their gathers - find it there
        const int lane = threadIdx.x & 31;
        for (int q = start + 32 * lane; q < end; q += 1024) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(colind + q));
            if constexpr (!UNIT) asm volatile("prefetch.global.L2 [%0];" ::"l"(val + q));
        }
    }
    const int mis = idx_mis;
    const int a0 = min(end, ((start + mis + 3) & ~3) - mis);     // first word boundary >= start
    const int a1 = max(a0, ((end + mis) & ~3) - mis);            // last word boundary <= end
    const int nw = (a1 - a0) >> 2;
    const int4 *cw = reinterpret_cast<const int4 *>(colind + a0);
    const T *vw = val + a0;

    // Register budget: the 4*NV gathers in flight (16*NV registers) must not be squeezed out by index storage, so the
    // index words live in ONE set of registers: as soon as the gathers of a round are issued the registers are
    // reloaded with the next round's indices - in the shadow of the gather latency (the index stream is an L2 hit
    // thanks to the prefetch above, so it arrives before the FMAs that wait for the gathers are done).
    int w = sub;
    int4 c[NV];
    Shfl v[NV][4];
#pragma unroll
    for (int n = 0; n < NV; ++n) {
        c[n] = make_int4(0, 0, 0, 0);
        if (w + n * P < nw) {
            c[n] = __ldcs(cw + w + n * P);
            if constexpr (!UNIT) ld_val4<T>(vw + 4 * (w + n * P), v[n]);
        }
    }
    // edges (their latency overlaps the first words' index loads)
    {
        const int nh = a0 - start, ne = nh + (end - a1);
        for (int e = sub; e < ne; e += P) csr_one<T, E, UNIT>(colind, val, ldb_bytes, e < nh ? start + e : a1 + (e - nh), Bcol, acc);
    }
    for (; w + (NV - 1) * P < nw; w += NV * P) {
        Pack<T, E> b[NV][4];
        Shfl vv[NV][4];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            b[n][0] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].x, ldb_bytes));
            b[n][1] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].y, ldb_bytes));
            b[n][2] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].z, ldb_bytes));
            b[n][3] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].w, ldb_bytes));
            if constexpr (!UNIT) {
#pragma unroll
                for (int k = 0; k < 4; ++k) vv[n][k] = v[n][k];
            }
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const int wn = w + (NV + n) * P;
            if (wn < nw) {
                c[n] = __ldcs(cw + wn);
                if constexpr (!UNIT) ld_val4<T>(vw + 4 * wn, v[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) {
#pragma unroll
            for (int k = 0; k < 4; ++k) fma_pack<T, E>(acc, b[n][k], UNIT ? (Shfl)1 : vv[n][k]);
        }
    }
    if constexpr (NV > 1) {
        // fewer than NV words left for this lane group: one at a time (c[0..] hold them in order)
#pragma unroll
        for (int n = 0; n < NV - 1; ++n) {
            if (w + n * P < nw) {
                Pack<T, E> b[4];
                b[0] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].x, ldb_bytes));
                b[1] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].y, ldb_bytes));
                b[2] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].z, ldb_bytes));
                b[3] = ld_dense<T, E>(row_ptr<T>(Bcol, c[n].w, ldb_bytes));
#pragma unroll
                for (int k = 0; k < 4; ++k) fma_pack<T, E>(acc, b[k], UNIT ? (Shfl)1 : v[n][k]);
            }
        }
    }
    return out;
}

// Alternative index delivery (the round-1 loop, kept selectable: -DPYGIM_IDX_SHFL=1 or NV == 0): the warp reads
// 32*R column ids + values per coalesced evict-first load, D batches ahead, and hands them round by shuffles.
template <typename T, int E, int G, int UNROLL, int R, int D, bool UNIT>
__device__ __forceinline__ void csr_accumulate_shfl(const CsrArgs<T> &a, int range_start, int range_end, const T *Bcol,
                                                    bool active, typename Arith<T>::Acc (&acc)[E]) {
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    constexpr int BATCH = 32 * R;
    constexpr int STEPS = G * R;                 // gather steps per full batch (P nonzeros each)
    constexpr int U = (UNROLL < STEPS) ? UNROLL : STEPS;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int end = range_end;
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
}

// One row (or one segment of a long row): accumulate, combine the lane groups, store / publish.
template <typename T, int E, int G, int NV, bool UNIT>
__device__ __forceinline__ void csr_process_range(const CsrArgs<T> &a, int range_start, int range_end, int chunk,
                                                  int dst_row, int long_idx) {
    using Acc = typename Arith<T>::Acc;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int vec = chunk * G + (lane % G);
    const bool active = vec < a.nvec;
    const T *Bcol = a.B + (long long)vec * E;
    asm volatile("" : "+l"(Bcol));     // keep the base in a register pair: gather address = one IMAD.WIDE.U32

    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    if constexpr (NV > 0) {
        const AccPack<T, E> sum = csr_accumulate<T, E, G, NV, UNIT>(a.colind, a.val, Bcol, a.ldb_bytes, a.idx_mis,
                                                                     range_start, range_end, active);
#pragma unroll
        for (int k = 0; k < E; ++k) acc[k] = sum.v[k];
    } else {
        constexpr int UNROLL = (E >= 8) ? 4 : 8;
        constexpr int R = (G >= UNROLL) ? 1 : ((UNROLL / G) > 4 ? 4 : (UNROLL / G));
        csr_accumulate_shfl<T, E, G, UNROLL, R, (R > 1) ? 1 : 2, UNIT>(a, range_start, range_end, Bcol, active, acc);
    }

    if (long_idx >= 0) {
#pragma unroll
        for (int off = G; off < 32; off <<= 1) {
#pragma unroll
            for (int k = 0; k < E; ++k) acc[k] += __shfl_xor_sync(FULL, acc[k], off);
        }
        if (sub == 0 && active)
            st_plain<T, E>(a.partial + (long long)dst_row * a.ldp + (long long)vec * E, narrow<T, E>(acc));
        csr_finish_long_row<T, E, G>(a, chunk, long_idx);     // rare path, kept out of line
        return;
    }
    csr_store_row<T, E, G>(a, acc, dst_row, vec, sub == 0 && active);
}

// SHORT-ROW graphs: the rows [ja, jb) of an item are consecutive, so their nonzeros are ONE contiguous stream.
// It is read in prefetched 32-entry batches that ignore row boundaries (the index-load latency of a row is hidden
// behind the rows before it) and walked run by run - a run being the part of a row inside the batch; row ends
// come from the item's rowptr entries held in the lanes (`rp`: lane l holds rowptr[first + l]).
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

// TINY ROWS (graphs of mean degree < ~12: citation networks): a launch of its own, one LANE GROUP per row.
// On arxiv-shape 92 % of the rows have at most 8 nonzeros; walking each of them with a whole warp costs a dependent
// chain (row bounds -> indices -> gathers -> shuffle tree -> store) of ~230 instructions and three memory latencies
// per ROW, and a persistent grid serialises those chains per warp (measured: 69 us per launch for 8 us of memory
// traffic).  Here the grid is as wide as the matrix: every group of G lanes owns one row, reads its (at most W)
// column indices at once (the same address across the group: one broadcast transaction), gathers one 16-byte word
// per lane and nonzero and stores its row: no tickets, no shuffles, ~3 instructions per
// nonzero, 64 resident warps per SM, every row of the matrix in flight within a few waves.  Rows longer than W are
// left to the persistent kernel (STREAM == 2 skips the rows taken here); the two launches write disjoint rows.
constexpr int kTinyRow = 8;       // longest row a lane group takes
template <typename T, int E, int G, bool UNIT>
__global__ void __launch_bounds__(256, (E >= 8) ? 4 : 8) csr_tiny_rows_kernel(const __grid_constant__ CsrArgs<T> a) {
    using Acc = typename Arith<T>::Acc;
    using Shfl = typename Arith<T>::Shfl;
    constexpr int W = kTinyRow;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (256 / G) + (int)threadIdx.x / G;
    const int vec = blockIdx.y * G + (lane % G);
    if (row >= a.nrows || vec >= a.nvec) return;
    const int i = __ldg(a.rowptr + row);
    const int n = __ldg(a.rowptr + row + 1) - i;
    if (n > W) return;
    const T *Bcol = a.B + (long long)vec * E;
    Acc acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
    // two halves of W/2 nonzeros: the second half's indices (rows of 5..8 nonzeros only) load while the first
    // half's gathers are in flight
    constexpr int HW = W / 2;
    int c[HW], c2[HW];
    Shfl v[HW], v2[HW];
#pragma unroll
    for (int u = 0; u < HW; ++u) {
        c[u] = 0;
        v[u] = (Shfl)1;
        if (u < n) {
            c[u] = ld_stream(a.colind + i + u);
            if constexpr (!UNIT) v[u] = ld_stream(a.val + i + u);
        }
    }
    Pack<T, E> b[HW];
#pragma unroll
    for (int u = 0; u < HW; ++u)
        if (u < n) b[u] = ld_dense<T, E>(row_ptr<T>(Bcol, c[u], a.ldb_bytes));
    if (n > HW) {
#pragma unroll
        for (int u = 0; u < HW; ++u) {
            c2[u] = 0;
            v2[u] = (Shfl)1;
            if (HW + u < n) {
                c2[u] = ld_stream(a.colind + i + HW + u);
                if constexpr (!UNIT) v2[u] = ld_stream(a.val + i + HW + u);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < HW; ++u)
        if (u < n) fma_pack<T, E>(acc, b[u], v[u]);
    if (n > HW) {
#pragma unroll
        for (int u = 0; u < HW; ++u)
            if (HW + u < n) b[u] = ld_dense<T, E>(row_ptr<T>(Bcol, c2[u], a.ldb_bytes));
#pragma unroll
        for (int u = 0; u < HW; ++u)
            if (HW + u < n) fma_pack<T, E>(acc, b[u], v2[u]);
    }
    csr_emit<T, E>(a, acc, row, vec);
}

// ---------------------------------------------------------------------------------------------- scheduling
struct CsrItem {
    int4 w;            // LOADED, and not touched before the item is processed (so the load stays asynchronous):
                       //   rows:    w.x = this lane's rowptr entry, lane l holds rowptr[first + l] (l <= count)
                       //   segment: the Seg descriptor (long_idx, start, end, slot)
};
// rows of row item `it` of superticket `sp` (arithmetic only)
__device__ __forceinline__ int csr_item_first(const int4 &sp, int it) { return sp.x + it * sp.z; }
__device__ __forceinline__ int csr_item_count(const int4 &sp, int it) { return min(sp.z, sp.y - it * sp.z); }

// Item `it` of superticket `sp`.  Its address follows from the ticket by arithmetic alone: ONE dependent load
// (rowptr entries, or the segment descriptor as one 16-byte word) between drawing a ticket and having the item.
// Nothing here consumes the loaded registers: any arithmetic on them would make the warp wait for the load right
// here instead of after the previous item's work (measured on arxiv-shape: two overlapping scalar loads into one
// register cost a full memory latency per item, 29 % of all stall samples).
// SEGS_ONLY (tiny-split plans: no row items): the row-pointer load is compiled out - predicated off it would still
// wait on the descriptor load's scoreboard (same destination register), 15 % of that kernel's stall samples.
template <typename T, bool SEGS_ONLY = false>
__device__ __forceinline__ CsrItem csr_load_item(const CsrArgs<T> &a, const int4 &sp, int it) {
    CsrItem r;
    const int lane = threadIdx.x & 31;
    if (SEGS_ONLY || sp.x < 0) r.w = __ldg(reinterpret_cast<const int4 *>(a.segs + (~sp.x + it)));
    else r.w.x = a.rowptr[min(csr_item_first(sp, it) + lane, a.nrows)];
    return r;
}

// Every warp calls this once, when it has found no more work: the last warp out zeroes the counters (the plan
// needs no host-side bookkeeping between launches; a launch can be replayed from a CUDA graph) and - in a
// row-sharded multi-GPU launch - tells every peer that this rank's rows have landed.
template <typename T> __device__ __forceinline__ void csr_leave(const CsrArgs<T> &a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    unsigned left = 0;
    if (lane == 0) {
        if (a.epi.n_peers > 0) __threadfence_system();      // this warp's peer / multimem stores before the count
        else asm volatile("fence.acq_rel.gpu;" ::: "memory");
        left = atomicAdd(a.warps_out, 1u);
    }
    left = __shfl_sync(FULL, left, 0);
    if (left != a.n_warps - 1u) return;
    const int n = a.n_super * a.col_chunks;
    for (int i = lane; i < n; i += 32) a.super_cnt[i] = 0;
    if (lane == 0) { a.warps_out[0] = 0u; a.warps_out[1] = 0u; }
    if (a.epi.n_peers > 0 && a.epi.flags[0] != nullptr) {
        __threadfence_system();           // every warp's rows (observed through the counter) before the flags
        if (lane < a.epi.n_peers)
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(a.epi.flags[lane] + a.epi.my_rank), "r"(a.epi.epoch)
                         : "memory");
    }
}

// Persistent, SM-affine grid.  (superticket s, column chunk c) has index s * col_chunks + c.
// THREADS only bounds the register allocation: the launcher picks the block size (256 .. THREADS).
template <typename T, int E, int G, int NV, int THREADS, int MIN_BLOCKS, bool UNIT, int STREAM = 0>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) csr_spmm_kernel(const __grid_constant__ CsrArgs<T> a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n_units = a.n_super * a.col_chunks;

    auto process = [&](const CsrItem &cur, const int4 &sp, int it, int chunk) {
        if (STREAM == 2 || sp.x < 0) {
            // one piece of a cut row (partial sum + merge by the last arriver), or - tiny-split plans - a whole row
            const int long_idx = cur.w.x, start = cur.w.y, end = cur.w.z;
            if (long_idx < 0) csr_process_range<T, E, G, NV, UNIT>(a, start, end, chunk, ~long_idx, -1);
            else csr_process_range<T, E, G, NV, UNIT>(a, start, end, chunk, cur.w.w, long_idx);
            return;
        }
        const int first = csr_item_first(sp, it), count = csr_item_count(sp, it);
        const int rp = cur.w.x;
        if (STREAM == 1) {
            // rows longer than seg_len are covered by their segments: stream the row blocks between them
            const int deg = __shfl_down_sync(FULL, rp, 1) - rp;
            const unsigned long_rows = __ballot_sync(FULL, lane < count && deg > a.seg_len);
            int ja = 0;
            while (ja < count) {
                const unsigned rest = long_rows >> ja;
                const int jb = rest ? ja + (__ffs(rest) - 1) : count;
                if (jb > ja) csr_stream_rows<T, E, G, 4, 2, UNIT>(a, first, ja, jb, rp, chunk);
                ja = jb + 1;
            }
        } else {
            // (STREAM == 2, the tiny-split family, has no row items: every row of its plans is a segment item)
            for (int j = 0; j < count; ++j) {
                const int start = __shfl_sync(FULL, rp, j);
                const int end = __shfl_sync(FULL, rp, j + 1);
                if (end - start <= a.seg_len)      // longer rows are covered by their segments
                    csr_process_range<T, E, G, NV, UNIT>(a, start, end, chunk, first + j, -1);
            }
        }
    };

    // drain one (superticket, chunk): one atomic per item.  The atomic of item k+2 is ISSUED while item k+1's
    // rowptr / descriptor load is issued and item k is processed; its result is only read an item later, so neither
    // the L2 atomic nor the dependent load is on the critical path.
    auto drain = [&](int unit, bool static_first) {
        const int s = unit / a.col_chunks, chunk = unit - s * a.col_chunks;
        const int4 sp = __ldg(a.supers + s);
        const int n = sp.w;
        int *cnt = a.super_cnt + unit;
        int pend = 0;                                   // lane 0: a ticket whose atomic may still be in flight
        int it;
        // `static_first` (the unit EVERY warp of the launch drains first): warp w starts on item w without asking -
        // otherwise the launch begins with n_warps atomics on one address, a few microseconds before the last warp
        // has its first item - and the counter hands out the tickets from n_warps on.
        const int base = static_first ? (int)a.n_warps : 0;
        if (static_first) {
            it = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
            if (lane == 0) pend = atomicAdd(cnt, 1) + base;
        } else {
            if (lane == 0) pend = atomicAdd(cnt, 1);
            it = __shfl_sync(FULL, pend, 0);
            if (lane == 0) {
                if (it == n) atomicAdd(a.warps_out + 1, 1u);      // exactly one warp draws ticket n: the unit is drawn out
                pend = atomicAdd(cnt, 1);
            }
        }
        CsrItem cur;
        if (it < n) cur = csr_load_item<T, STREAM == 2>(a, sp, it);
        while (it < n) {
            const int nit = __shfl_sync(FULL, pend, 0);          // drawn one item ago
            if (lane == 0 && nit == n) atomicAdd(a.warps_out + 1, 1u);
            CsrItem nxt;
            if (nit < n) {
                if (lane == 0) pend = atomicAdd(cnt, 1) + base;  // for the item after next: not awaited here
                nxt = csr_load_item<T, STREAM == 2>(a, sp, nit);
            }
            process(cur, sp, it, chunk);
            it = nit;
            cur = nxt;
        }
    };

    // Home supertickets are keyed by the SM, not by the block, so every block resident on an SM drains the same
    // ones; when they are exhausted the warp steals from any (superticket, chunk) that still has undrawn items,
    // scanning the draw counters 32 at a time from an SM-specific offset.
    unsigned smid, nsmid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsmid));
    // the block's warps share how far the SM's home list is already drawn out: a warp that finds a unit exhausted
    // publishes it, the others skip straight past instead of paying an L2 round trip per exhausted unit
    __shared__ int s_home_done;
    if (threadIdx.x == 0) s_home_done = 0;
    __syncthreads();
    int home_k = 0, scan_o = 0, scan_u = 0;           // home unit = smid + home_k * nsmid
    unsigned scan_m = 0;
    // ONE QUEUE (the default plan: no locality to exploit, n_units = segments + rows per column chunk): every warp
    // drains the units in order - exactly balanced, one atomic per item, nothing to scan.  SM-affine home lists with
    // stealing are for plans with a locality-preserving row order (row map set), where they pay.
    const bool one_queue = n_units <= 8;
    for (;;) {
        int unit;
        if (one_queue) {
            if (home_k >= n_units) break;
            drain(home_k, home_k == 0);
            ++home_k;
            continue;
        }
        {   // one lane reads the shared progress (lanes need not be converged here), all lanes take its value
            int k = home_k;
            if (lane == 0) k = max(k, *(volatile int *)&s_home_done);
            home_k = __shfl_sync(FULL, k, 0);
        }
        const bool is_home = (int)smid + home_k * (int)nsmid < n_units;
        if (is_home) {
            unit = (int)smid + home_k * (int)nsmid;
        } else {
            // scan the draw counters, 32 at a time, for units that still have undrawn items
            while (scan_m == 0 && scan_o < n_units) {
                // nothing left to draw anywhere?  (read by ONE lane, as above)
                unsigned drawn = 0;
                if (lane == 0) drawn = *(volatile unsigned int *)(a.warps_out + 1);
                if (__shfl_sync(FULL, drawn, 0) >= (unsigned)n_units) {
                    scan_o = n_units;
                    break;
                }
                const int base = (int)(((unsigned long long)smid * 2654435761ull) % (unsigned)n_units);
                scan_u = base + scan_o + lane;
                if (scan_u >= n_units) scan_u -= n_units;
                bool open = false;
                if (scan_o + lane < n_units)
                    open = *(volatile int *)(a.super_cnt + scan_u) < __ldg(&a.supers[scan_u / a.col_chunks].w);
                scan_m = __ballot_sync(FULL, open);
                scan_o += 32;
            }
            if (scan_m == 0) break;
            const int src = __ffs(scan_m) - 1;
            scan_m &= scan_m - 1;
            unit = __shfl_sync(FULL, scan_u, src);
        }
        drain(unit, false);
        if (is_home) {          // drain returns only when the unit is drawn out: the block's other warps can skip it
            if (lane == 0) atomicMax(&s_home_done, home_k + 1);
            ++home_k;
        }
    }
    csr_leave<T>(a);
}

}  // namespace pygim
