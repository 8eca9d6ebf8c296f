// Hot/cold CSR SpMM for graphs with community structure: the dense rows a superticket gathers most are staged in
// SHARED MEMORY once and gathered from there; only the rest crosses L2 -> SM.
//
// Why: on Reddit-like graphs the plain kernel (spmm_csr.cuh) is bound by L2 -> SM traffic - every gathered byte
// crosses the crossbar (profiles/r02_*).  A row order that puts rows sharing neighbours next to each other
// (pygim_b200/reorder.py) makes reuse POSSIBLE, but the hardware L1 keeps less than half of a community's rows
// while a third of the traffic streams through it (measured: 33 % sector hits at H = 32).  A software-managed tile
// keeps exactly the rows that pay: the plan (reorder.py::hot_cold_plan -> pygim_plan_set_hot_tiles) picks, per
// superticket, the K most referenced columns, stores each row's nonzeros hot-first (hot entries hold a TILE SLOT
// instead of a column id) and records the per-row hot count.
//
// Kernel: one block per SM, the blocks draw (superticket, 128-byte column chunk) units from a global counter;
//   1. the block loads the unit's K hot dense rows (this chunk of them) into shared memory - coalesced, once;
//   2. its warps draw the unit's row items from a shared-memory counter; a row's hot nonzeros are gathered with
//      LDS.128 out of the tile, its cold ones with LDG.128 as in the plain kernel; same vector index loads, same
//      fixed-order shuffle tree, same row store (csr_emit: row map, de-quantise, residual, peers);
//   3. segments of long rows (all cold) are units of their own and use no tile.
// Everything a row sums is the same set of products as in the plain kernel - integers and integer-valued floats are
// bit-identical, real-valued floats differ only by summation order (hot first).
#pragma once
#include "spmm_csr.cuh"

namespace pygim {

template <typename T> struct HcArgs {
    CsrArgs<T> c;            // the plain arguments (supers, segs, epilogue ...); c.super_cnt[0] is the unit counter
    const int *hot_cols;     // [n_row_supers x hot_k] column of every tile slot, -1 = unused
    const int *hot_cnt;      // [nrows] leading nonzeros of the row that are tile slots
    int hot_k;               // tile rows
    int n_seg_super;         // supers[0 .. n_seg_super) are segment supertickets (no tile)
};

// acc += sum over the nonzeros [start, end) - all HOT: colind holds tile slots - of val * tile[slot, lane word]
template <typename T, int E, int G, int NV, bool UNIT>
__device__ __forceinline__ void hc_accumulate_tile(const int *colind, const T *val, const uint4 *tile_lane, int idx_mis,
                                                   int start, int end, typename Arith<T>::Acc (&acc)[E]) {
    using Shfl = typename Arith<T>::Shfl;
    constexpr int P = 32 / G;
    const int sub = (threadIdx.x & 31) / G;
    auto lds = [&](int slot) -> Pack<T, E> {
        union { uint4 w; Pack<T, E> v; } u;
        u.w = tile_lane[slot * G];
        return u.v;
    };
    auto one = [&](int i) {
        const int slot = __ldg(colind + i);
        Shfl v = (Shfl)1;
        if constexpr (!UNIT) v = (Shfl)__ldg(val + i);
        fma_pack<T, E>(acc, lds(slot), v);
    };
    if (idx_mis >= 4) {
        for (int i = start + sub; i < end; i += P) one(i);
        return;
    }
    const int mis = idx_mis;
    const int a0 = min(end, ((start + mis + 3) & ~3) - mis);
    const int a1 = max(a0, ((end + mis) & ~3) - mis);
    const int nw = (a1 - a0) >> 2;
    const int4 *cw = reinterpret_cast<const int4 *>(colind + a0);
    const T *vw = val + a0;
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
    {
        const int nh = a0 - start, ne = nh + (end - a1);
        for (int e = sub; e < ne; e += P) one(e < nh ? start + e : a1 + (e - nh));
    }
    for (; w < nw; w += NV * P) {
        int4 cc[NV];
        Shfl vv[NV][4];
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            cc[n] = c[n];
            if constexpr (!UNIT) {
#pragma unroll
                for (int k = 0; k < 4; ++k) vv[n][k] = v[n][k];
            }
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) {        // the next round's indices load while this round reads the tile
            const int wn = w + (NV + n) * P;
            if (wn < nw) {
                c[n] = __ldcs(cw + wn);
                if constexpr (!UNIT) ld_val4<T>(vw + 4 * wn, v[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            if (w + n * P < nw) {
                const Pack<T, E> b0 = lds(cc[n].x), b1 = lds(cc[n].y), b2 = lds(cc[n].z), b3 = lds(cc[n].w);
                fma_pack<T, E>(acc, b0, UNIT ? (Shfl)1 : vv[n][0]);
                fma_pack<T, E>(acc, b1, UNIT ? (Shfl)1 : vv[n][1]);
                fma_pack<T, E>(acc, b2, UNIT ? (Shfl)1 : vv[n][2]);
                fma_pack<T, E>(acc, b3, UNIT ? (Shfl)1 : vv[n][3]);
            }
        }
    }
}

// One block per SM; THREADS only bounds the registers.  Dynamic shared memory: hot_k x G x 16 bytes.
template <typename T, int E, int G, int NV, int THREADS, bool UNIT>
__global__ void __launch_bounds__(THREADS, 1) csr_hc_kernel(const __grid_constant__ HcArgs<T> h) {
    using Acc = typename Arith<T>::Acc;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ uint4 hc_tile[];
    __shared__ int s_unit, s_next;
    const CsrArgs<T> &a = h.c;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int n_units = a.n_super * a.col_chunks;

    for (;;) {
        __syncthreads();                                   // the previous unit's tile and counters are free
        if (threadIdx.x == 0) {
            s_unit = atomicAdd(a.super_cnt, 1);
            s_next = 0;
        }
        __syncthreads();
        const int unit = s_unit;
        if (unit >= n_units) break;
        const int s = unit / a.col_chunks, chunk = unit - s * a.col_chunks;
        const int4 sp = __ldg(a.supers + s);
        const int vec = chunk * G + (lane % G);
        const bool active = vec < a.nvec;
        if (sp.x >= 0) {
            // stage this chunk of the unit's hot dense rows: G consecutive threads copy one 16*G-byte row piece
            const int *cols = h.hot_cols + (long long)(s - h.n_seg_super) * h.hot_k;
            const int total = h.hot_k * G;
            for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
                int col[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * blockDim.x;
                    col[u] = i < total ? __ldg(cols + i / G) : -1;
                }
                uint4 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * blockDim.x;
                    const int v = chunk * G + (i % G);
                    w[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (col[u] >= 0 && v < a.nvec)
                        w[u] = __ldg(reinterpret_cast<const uint4 *>(row_ptr<T>(a.B + (long long)v * E, col[u], a.ldb_bytes)));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * blockDim.x;
                    if (i < total) hc_tile[i] = w[u];
                }
            }
        }
        __syncthreads();
        const uint4 *tile_lane = hc_tile + (lane % G);
        const T *Bcol = a.B + (long long)vec * E;
        asm volatile("" : "+l"(Bcol));

        // row items (or segments) of the unit, drawn from the shared-memory counter; the next item's rowptr /
        // hot-count loads are issued before the current item is processed
        auto take = [&]() -> int {
            int t = 0;
            if (lane == 0) t = atomicAdd(&s_next, 1);
            return __shfl_sync(FULL, t, 0);
        };
        int it = take();
        CsrItem cur;
        int cur_hot = 0;
        if (it < sp.w) {
            cur = csr_load_item<T>(a, sp, it);
            if (sp.x >= 0) cur_hot = __ldg(h.hot_cnt + min(csr_item_first(sp, it) + lane, a.nrows - 1));
        }
        while (it < sp.w) {
            const int nit = take();
            CsrItem nxt;
            int nxt_hot = 0;
            if (nit < sp.w) {
                nxt = csr_load_item<T>(a, sp, nit);
                if (sp.x >= 0) nxt_hot = __ldg(h.hot_cnt + min(csr_item_first(sp, nit) + lane, a.nrows - 1));
            }
            if (sp.x < 0) {
                csr_process_range<T, E, G, NV, UNIT>(a, cur.w.y, cur.w.z, chunk, cur.w.w, cur.w.x);
            } else {
                const int first = csr_item_first(sp, it), count = csr_item_count(sp, it);
                for (int j = 0; j < count; ++j) {
                    const int start = __shfl_sync(FULL, cur.w.x, j);
                    const int end = __shfl_sync(FULL, cur.w.x, j + 1);
                    const int hot = __shfl_sync(FULL, cur_hot, j);
                    if (end - start > a.seg_len) continue;          // covered by its (cold) segments
                    Acc acc[E];
#pragma unroll
                    for (int k = 0; k < E; ++k) acc[k] = (Acc)0;
                    // the whole row's index (and value) stream into L2 now: the tile gathers have no latency of their
                    // own to hide behind, so the index words must already be close
                    for (int q = start + 32 * lane; q < end; q += 1024) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.colind + q));
                        if constexpr (!UNIT) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.val + q));
                    }
                    if (active)
                        hc_accumulate_tile<T, E, G, (UNIT && E < 8) ? 8 : 2, UNIT>(a.colind, a.val, tile_lane, a.idx_mis, start,
                                                                                  start + hot, acc);
                    const AccPack<T, E> cold = csr_accumulate<T, E, G, NV, UNIT>(a.colind, a.val, Bcol, a.ldb_bytes, a.idx_mis,
                                                                                 start + hot, end, active);
#pragma unroll
                    for (int k = 0; k < E; ++k) acc[k] += cold.v[k];
                    csr_store_row<T, E, G>(a, acc, first + j, vec, sub == 0 && active);
                }
            }
            it = nit;
            cur = nxt;
            cur_hot = nxt_hot;
        }
    }
    csr_leave<T>(a);
}

}  // namespace pygim
