// libbackend_pim.so - C ABI (include/pygim_b200.h), device context and plans.
//
// Host-side counterpart of the reference's op layer (spmm_default/pytorch_api.cpp, ops.hpp,
// spmm_mul_csr.c, spmm_mul_coo.c): bring-up, one-time sparse upload + partition plan
// ("to_device_group"), and the per-call run ("run_group").  There is no CPU fallback: without a
// CUDA device every entry point that would compute returns PYGIM_ERR_NO_DEVICE.
#include "../../include/pygim_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "launch.h"
#include "spmm_csr.cuh"   // struct Seg

namespace pygim {

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(PYGIM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

static size_t dtype_size(int dt) {
    switch (dt) {
        case PYGIM_INT8: return 1;
        case PYGIM_INT16: return 2;
        case PYGIM_INT32: return 4;
        case PYGIM_INT64: return 8;
        case PYGIM_FLT32: return 4;
        case PYGIM_DBL64: return 8;
        default: return 0;
    }
}

// ------------------------------------------------------------------------------ device context
struct Context {
    bool initialised = false;
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    int max_threads_per_sm = 2048;
    long long l2_bytes = 0, persisting_max = 0, hbm_bytes = 0;
    long long persisting_set = 0, max_window = 0;   // L2 set aside for persisting lines / largest policy window
    long long nr_ranks = 0, groups_per_rank = 1, nr_dpus = 0;
    std::mutex mu;
};
static Context g_ctx;

static int context_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return fail(PYGIM_ERR_NO_DEVICE,
                    "no CUDA device visible (%s); pygim_b200 has no CPU fallback", cudaGetErrorString(e));
    }
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) return fail(PYGIM_ERR_INVALID, "device %d out of range (%d visible)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    g_ctx.device = device;
    g_ctx.sm_count = p.multiProcessorCount;
    g_ctx.cc_major = p.major;
    g_ctx.cc_minor = p.minor;
    g_ctx.max_threads_per_sm = p.maxThreadsPerMultiProcessor;
    g_ctx.l2_bytes = p.l2CacheSize;
    g_ctx.persisting_max = p.persistingL2CacheMaxSize;
    g_ctx.hbm_bytes = (long long)p.totalGlobalMem;
    g_ctx.max_window = p.accessPolicyMaxWindowSize;
    // set the persisting-L2 carve-out aside once; plans opt in per launch ("l2_persist" option)
    g_ctx.persisting_set = 0;
    if (p.persistingL2CacheMaxSize > 0 &&
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)p.persistingL2CacheMaxSize) == cudaSuccess) {
        size_t got = 0;
        if (cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize) == cudaSuccess) g_ctx.persisting_set = (long long)got;
    }
    (void)cudaGetLastError();
    g_ctx.initialised = true;
    return PYGIM_OK;
}

// ------------------------------------------------------------------------------ plans
// Second-level balancing of a CSR row range, built on the host from rowptr: work ITEMS in row order (a group of
// consecutive short rows, or one seg_len-bounded segment of a long row) bundled into SUPERTICKETS of near-equal
// nnz (spmm_csr.cuh).  `full` covers every row; the host entry point additionally keeps a few nnz-balanced row
// chunks so that the download of chunk k overlaps the kernel of chunk k+1.
struct CsrPlan {
    long long row_begin = 0, row_end = 0;
    Seg *d_segs = nullptr;
    int4 *d_supers = nullptr;
    int *d_long_rows = nullptr;
    int *d_long_seg_ptr = nullptr;
    int n_items = 0, n_super = 0, n_seg_super = 0;
    int n_seg = 0, n_long = 0;
    int n_slots = 0;          // rows of the partial-sum scratch (pieces of rows cut into several segments)
    bool tiny_split = false;  // two-launch family: rows of <= kTinyRow nonzeros are NOT in this plan (csr_tiny_rows_kernel
                              // takes them), every other row is a list of seg_len-bounded pieces, longest first
};

struct SparsePart {
    long long nrows = 0, ncols = 0, nnz = 0;
    const int *rowidx = nullptr;   // CSR rowptr [nrows+1] or COO rowind [nnz]
    const int *colind = nullptr;
    const void *values = nullptr;
    bool owned = false;            // true: uploaded by us, freed in free_group
    bool unit_values = false;      // every stored value == 1 (checked on the device at plan time)
    // COO: the stream is row-major sorted (checked on the device at plan time).  A sorted stream is run through
    // the CSR kernels over a row pointer derived here once (d_rowptr); an unsorted one through the COO kernel with
    // every flush atomic.
    bool coo_sorted = true;
    int *d_rowptr = nullptr;       // COO only: derived row pointer [nrows + 1] (owned)
    // CSR second-level balancing (built from rowptr on the host)
    std::vector<int> h_rowptr;     // host copy (kept for re-planning when seg_len changes)
    CsrPlan full;
    std::vector<CsrPlan> chunks;   // lazily built by the host entry point
    int seg_len = 0;
    long long max_row_nnz = 0, empty_rows = 0;
    // hot/cold plan (pygim_plan_set_hot_tiles): row supertickets at the caller's boundaries, each with hot_k tile
    // columns; colind then holds tile slots for the first hot_cnt[r] nonzeros of row r
    std::vector<int> hot_super_rows;
    int *d_hot_cols = nullptr, *d_hot_cnt = nullptr;
    int hot_k = 0;
    const int *csr_rowptr() const { return d_rowptr ? d_rowptr : rowidx; }
};

// Mutable launch state.  One per (plan, stream): launches on one stream are serialised, launches on different
// streams get their own counters and partial-sum scratch, so one handle may be used from several streams.
// Every counter is zero at rest (the last warp of a launch resets them), so the buffer can be shared by all
// row-range plans of the group.
struct Scratch {
    cudaStream_t stream = nullptr;
    int *counters = nullptr;        // [0] warps out, [4 ..) superticket draw counters, then long-row arrival counters
    size_t n_counters = 0;
    void *partial = nullptr;
    size_t partial_bytes = 0;
    unsigned long long *coo_ticket = nullptr;   // COO kernel: work counter + warps out
};

struct Group {
    int format = PYGIM_CSR;
    int dtype = PYGIM_FLT32;
    int device = 0;
    bool csr_view = false;          // COO plan whose parts are all sorted: runs through the CSR kernels
    long long h_size = 0, total_rows = 0, total_cols = 0;
    std::vector<SparsePart> parts;
    std::vector<long long> dense_cols;
    // options (< 0 = automatic)
    long long opt_seg_len = -1, opt_l2_persist = -1, opt_chunk_nnz = -1, opt_rows_per_ticket = -1;
    long long opt_item_nnz = -1, opt_super_nnz = -1, opt_max_g = -1, opt_cta_threads = -1;
    long long opt_unit_values = -1;   // 0 forces the general (weighted) kernels
    long long opt_short_rows = -1;    // 1/0 force the high-occupancy / deep-unroll CSR instantiation
    long long opt_host_chunks = -1;   // host entry point: row chunks for download/compute overlap (0 = off)
    long long opt_host_tile_bytes = -1;   // host entry point: bytes per row of one upload/compute column tile (default 256)
    long long opt_coo_native = -1;    // 1: sorted COO streams also run through the COO kernel
    int *d_row_map = nullptr;         // plan row r -> result row (row reordering); null = identity
    std::mutex mu;                    // guards `scratch`
    std::vector<Scratch> scratch;
    void *d_B = nullptr;   // staging for the host entry point
    void *d_C = nullptr;
    size_t dB_bytes = 0, dC_bytes = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // host entry point: row chunks + a second stream so the download of chunk k overlaps the kernel of chunk k+1
    cudaStream_t copy_stream = nullptr;      // downloads
    cudaStream_t copy_in_stream = nullptr;   // uploads
    std::vector<cudaEvent_t> chunk_done;
    double timers_ms[5] = {0, 0, 0, 0, 0};
    int64_t last_launches = 0;
};

// Kernel family of a sparse part: 0 deep (128 registers, 16 gathers in flight) for long rows; 3 light (64 registers,
// twice the warps) for short rows; 4 the two-launch family for very short rows (citation graphs: most rows have a
// handful of nonzeros and go to csr_tiny_rows_kernel); 1 high occupancy and 2 streamed row items stay selectable.
static int kernel_family(const Group &g, const SparsePart &p) {
    if (g.opt_short_rows >= 0) return (int)g.opt_short_rows;
    const long long rows1 = std::max<long long>(p.nrows, 1);
    return p.nnz < 12 * rows1 ? 4 : (p.nnz < 96 * rows1 ? 3 : 0);
}

static int auto_seg_len(const Group &g, const SparsePart &p, long long row_bytes) {
    // A segment costs a release fence, a partial-sum round trip and (for the last arriver) an acquire, so segments
    // should be as long as balance allows: ~6 items per resident warp of the deep kernel family (16 warps per SM),
    // twice that for narrow dense rows (<= 128 bytes), 512..4096 nonzeros.  Measured (FLT32, H = 16/32/64/128):
    // whole Reddit-shape 4096 best (9630 GFLOP/s vs 9390 at 1024); a 1/8 row shard 2048/2048/1024/1024 best
    // (81/113/208/500 us vs 102/130/225/513 at 512 and 146/169/257/542 at 256).
    // Short-row graphs run the light family (32 warps per SM) and have few long rows: there balance wins
    // (products-shape sweep 2670 GFLOP/s at 1024 vs 2510 at 4096).
    // The two-launch family is latency-bound (a launch is a few waves of dependent chains): no piece longer than
    // four gather rounds of one warp.
    if (kernel_family(g, p) == 4 && p.hot_super_rows.empty()) return 256;
    const bool short_rows = p.nnz < 96 * std::max<long long>(p.nrows, 1);
    const long long slots = (long long)g_ctx.sm_count * (short_rows ? 64 : 16);
    long long s = p.nnz / std::max<long long>(1, slots * (short_rows ? 8 : 6));
    if (!short_rows && row_bytes > 0 && row_bytes <= 128) s *= 2;
    long long pow2 = 512;
    while (pow2 < s && pow2 < 4096) pow2 <<= 1;
    return (int)pow2;
}

static void free_plan(CsrPlan &c) {
    if (c.d_segs) cudaFree(c.d_segs);
    if (c.d_supers) cudaFree(c.d_supers);
    if (c.d_long_rows) cudaFree(c.d_long_rows);
    if (c.d_long_seg_ptr) cudaFree(c.d_long_seg_ptr);
    c = CsrPlan();
}

static void free_csr_plan(SparsePart &p) {
    free_plan(p.full);
    for (auto &c : p.chunks) free_plan(c);
    p.chunks.clear();
}

// bytes of the widest dense tile row a launch of this group gathers
static long long widest_tile_bytes(const Group &g) {
    long long w = 0;
    for (long long c : g.dense_cols) w = std::max(w, c);
    return w * (long long)dtype_size(g.dtype);
}

template <typename V> static int upload(V **dst, const std::vector<V> &src) {
    *dst = nullptr;
    if (src.empty()) return PYGIM_OK;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(dst), src.size() * sizeof(V)));
    CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(V), cudaMemcpyHostToDevice));
    return PYGIM_OK;
}

// Supertickets of the rows [r0, r1).  A row longer than seg_len becomes ceil(nnz/seg_len) near-equal segments (their
// partial sums are merged in the kernel); the segments form the first supertickets, longest first.  The rows
// themselves are cut, in row order, into supertickets of about super_nnz nonzeros; inside one the rows are dealt
// `rows_per_item` at a time (about item_nnz nonzeros, at most 31 rows) - an item's rows follow from its ticket by
// arithmetic, no descriptor load.  Row ids are relative to r0 (the kernel is handed rowptr + r0), nonzero offsets
// stay absolute.
static int build_plan_range(const Group &g, const SparsePart &p, int seg_len, long long r0, long long r1, CsrPlan &out) {
    free_plan(out);
    out.row_begin = r0;
    out.row_end = r1;
    std::vector<Seg> segs;
    std::vector<int> long_rows, long_ptr;
    long_ptr.push_back(0);
    const std::vector<int> &rp = p.h_rowptr;
    // ~256 nonzeros per item, fewer on small graphs so that every resident warp still gets about four items
    long long target = g.opt_item_nnz;
    if (target <= 0) {
        const long long want = p.nnz / std::max<long long>(1, 4LL * g_ctx.sm_count * 32);
        target = 64;
        while (target * 2 <= want && target < 256) target *= 2;
    }
    const int max_rows = (int)std::max<long long>(1, std::min<long long>(31, g.opt_rows_per_ticket > 0 ? g.opt_rows_per_ticket : 31));
    long long short_nnz = 0;
    const bool tiny_split = kernel_family(g, p) == 4 && p.hot_super_rows.empty();
    out.tiny_split = tiny_split;
    int n_slots = 0;
    for (long long r = r0; r < r1; ++r) {
        const long long s = (unsigned)rp[r], e = (unsigned)rp[r + 1];
        const long long n = e - s;
        if (tiny_split && (n <= seg_len || n <= (long long)pygim::kTinyRow)) {
            // rows of at most kTinyRow nonzeros belong to csr_tiny_rows_kernel (whatever seg_len says); every other
            // uncut row is ONE piece that stores its row directly (long_idx = ~row: no partial sum, no merge)
            if (n > (long long)pygim::kTinyRow) {
                Seg sg;
                sg.long_idx = ~(int)(r - r0);
                sg.start = (int)s;
                sg.end = (int)e;
                sg.slot = 0;
                segs.push_back(sg);
            }
            continue;
        }
        if (n > seg_len) {
            const long long k = (n + seg_len - 1) / seg_len;
            // equal pieces rounded up to a multiple of 32 so every piece but the last runs full rounds
            const long long piece = ((n + k - 1) / k + 31) / 32 * 32;
            for (long long b = s; b < e; b += piece) {
                Seg sg;
                sg.long_idx = (int)long_rows.size();
                sg.start = (int)b;
                sg.end = (int)std::min(e, b + piece);
                sg.slot = n_slots++;
                segs.push_back(sg);
            }
            long_rows.push_back((int)(r - r0));
            long_ptr.push_back(n_slots);
        } else {
            short_nnz += n;
        }
    }
    long long seg_nnz = 0;
    for (const Seg &sg : segs) seg_nnz += sg.end - sg.start;
    // Supertickets.  Default: ONE for the segments and ONE for the rows - a single shared queue each, exactly
    // balanced (measured on a 1/8 Reddit-shape shard and on arxiv-shape: SM-affine lists with stealing cost 50-80 us
    // of scanning and imbalance per launch, and buy nothing on a graph in arbitrary order).  With a row map (the rows
    // were reordered for locality) or an explicit super_nnz: ~8 per SM, at least 2048 nonzeros, at most 256 K.
    long long super = g.opt_super_nnz;
    if (super <= 0) {
        // (measured on the clustered Reddit-shape graph: the SM-affine schedule wins 2-4 % with dense rows of 256 bytes
        // and more and loses 12-24 % below - 64 / 128-byte rows are bound by L1 tag lookups, not by the crossbar)
        if ((g.d_row_map != nullptr && widest_tile_bytes(g) >= 256) || !p.hot_super_rows.empty())
            super = std::min<long long>(262144, std::max<long long>(2048, (short_nnz + seg_nnz) / std::max(1, g_ctx.sm_count * 8)));
        else
            super = (1LL << 60);
    }
    std::vector<int4> supers;
    long long n_items = 0;
    if (!segs.empty()) {
        // longest pieces first: tickets are handed out in index order (slots keep the row order for the merge)
        std::stable_sort(segs.begin(), segs.end(),
                         [](const Seg &x, const Seg &y) { return (x.end - x.start) > (y.end - y.start); });
        long long acc = 0;
        int first = 0;
        for (size_t k = 0; k < segs.size(); ++k) {
            acc += segs[k].end - segs[k].start;
            if (acc >= super || k + 1 == segs.size()) {
                supers.push_back(make_int4(~first, 0, 0, (int)k + 1 - first));
                n_items += (long long)k + 1 - first;
                first = (int)k + 1;
                acc = 0;
            }
        }
    }
    out.n_seg_super = (int)supers.size();
    if (!tiny_split) {
        long long acc = 0, first = r0;
        size_t next_cut = 1;               // hot/cold plans: the caller's superticket boundaries
        const bool fixed = !p.hot_super_rows.empty() && r0 == 0 && r1 == p.nrows;
        for (long long r = r0; r < r1; ++r) {
            const long long n = (long long)(unsigned)rp[r + 1] - (long long)(unsigned)rp[r];
            acc += n > seg_len ? 0 : std::max<long long>(n, 1);
            bool cut = acc >= super || r + 1 == r1;
            if (fixed) {
                cut = (long long)p.hot_super_rows[next_cut] == r + 1;
                if (cut) ++next_cut;
            }
            if (cut) {
                const long long rows = r + 1 - first;
                const long long per = std::max<long long>(1, std::min<long long>(max_rows, (target * rows + acc / 2) / std::max<long long>(acc, 1)));
                const long long items = (rows + per - 1) / per;
                supers.push_back(make_int4((int)(first - r0), (int)rows, (int)per, (int)items));
                n_items += items;
                first = r + 1;
                acc = 0;
            }
        }
    }
    out.n_items = (int)std::min<long long>(n_items, 0x7fffffff);
    out.n_super = (int)supers.size();
    out.n_seg = (int)segs.size();
    out.n_slots = n_slots;
    out.n_long = (int)long_rows.size();
    int rc;
    if ((rc = upload(&out.d_supers, supers))) return rc;
    if (out.n_seg > 0) {
        if ((rc = upload(&out.d_segs, segs)) || (rc = upload(&out.d_long_rows, long_rows)) ||
            (rc = upload(&out.d_long_seg_ptr, long_ptr)))
            return rc;
    }
    return PYGIM_OK;
}

static int build_csr_plan(const Group &g, SparsePart &p, int seg_len) {
    free_csr_plan(p);
    p.seg_len = seg_len;
    p.max_row_nnz = 0;
    p.empty_rows = 0;
    for (long long r = 0; r < p.nrows; ++r) {
        const long long n = (long long)(unsigned)p.h_rowptr[r + 1] - (long long)(unsigned)p.h_rowptr[r];
        if (n > p.max_row_nnz) p.max_row_nnz = n;
        if (n == 0) ++p.empty_rows;
    }
    return build_plan_range(g, p, seg_len, 0, p.nrows, p.full);
}

static int replan(Group *g) {
    for (auto &p : g->parts) {
        if (p.h_rowptr.empty()) continue;
        int rc = build_csr_plan(*g, p, g->opt_seg_len > 0 ? (int)std::min<long long>(g->opt_seg_len, 1 << 30)
                                                        : auto_seg_len(*g, p, widest_tile_bytes(*g)));
        if (rc) return rc;
    }
    return PYGIM_OK;
}

// Handles are never-reused ids into a registry (a freed handle can only ever be "unknown", never somebody
// else's plan - the reference hands out raw pointers, pytorch_api.cpp:240).
static std::mutex g_reg_mu;
static std::unordered_map<uint64_t, Group *> g_registry;
static uint64_t g_next_handle = 1;

static Group *as_group(pygim_handle_t h) {
    std::lock_guard<std::mutex> lock(g_reg_mu);
    auto it = g_registry.find(h);
    if (it == g_registry.end()) {
        fail(PYGIM_ERR_INVALID, "unknown or freed plan handle %llu", (unsigned long long)h);
        return nullptr;
    }
    return it->second;
}

static void destroy_group(Group *g) {
    if (!g) return;
    for (auto &p : g->parts) {
        free_csr_plan(p);
        if (p.d_rowptr) cudaFree(p.d_rowptr);
        if (p.d_hot_cols) cudaFree(p.d_hot_cols);
        if (p.d_hot_cnt) cudaFree(p.d_hot_cnt);
        if (p.owned) {
            cudaFree(const_cast<int *>(p.rowidx));
            cudaFree(const_cast<int *>(p.colind));
            cudaFree(const_cast<void *>(p.values));
        }
    }
    for (auto &sc : g->scratch) {
        if (sc.counters) cudaFree(sc.counters);
        if (sc.partial) cudaFree(sc.partial);
        if (sc.coo_ticket) cudaFree(sc.coo_ticket);
    }
    if (g->d_row_map) cudaFree(g->d_row_map);
    if (g->d_B) cudaFree(g->d_B);
    if (g->d_C) cudaFree(g->d_C);
    for (auto &e : g->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : g->chunk_done)
        if (e) cudaEventDestroy(e);
    if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
    if (g->copy_in_stream) cudaStreamDestroy(g->copy_in_stream);
    delete g;
}

// The scratch of `stream` with at least n_counters zeroed ints and partial_bytes of partial-sum space.
static int get_scratch(Group *g, cudaStream_t stream, size_t n_counters, size_t partial_bytes, bool coo, Scratch *out) {
    std::lock_guard<std::mutex> lock(g->mu);
    Scratch *sc = nullptr;
    for (auto &c : g->scratch)
        if (c.stream == stream) sc = &c;
    if (!sc) {
        g->scratch.emplace_back();
        sc = &g->scratch.back();
        sc->stream = stream;
    }
    if (n_counters > sc->n_counters) {
        // grow-only; stream-ordered free keeps earlier launches of this stream valid
        if (sc->counters) CUDA_TRY(cudaFreeAsync(sc->counters, stream));
        const size_t n = std::max<size_t>(n_counters + n_counters / 2, 1024);
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&sc->counters), n * sizeof(int), stream));
        CUDA_TRY(cudaMemsetAsync(sc->counters, 0, n * sizeof(int), stream));
        sc->n_counters = n;
    }
    if (partial_bytes > sc->partial_bytes) {
        if (sc->partial) CUDA_TRY(cudaFreeAsync(sc->partial, stream));
        CUDA_TRY(cudaMallocAsync(&sc->partial, partial_bytes, stream));
        sc->partial_bytes = partial_bytes;
    }
    if (coo && !sc->coo_ticket) {
        CUDA_TRY(cudaMallocAsync(reinterpret_cast<void **>(&sc->coo_ticket), 2 * sizeof(unsigned long long), stream));
        CUDA_TRY(cudaMemsetAsync(sc->coo_ticket, 0, 2 * sizeof(unsigned long long), stream));
    }
    *out = *sc;       // a copy: the vector may grow under another thread
    return PYGIM_OK;
}

// ---- sorted COO -> row pointer (type independent): rowptr[r] = first nonzero whose row is >= r
__global__ void coo_check_sorted_kernel(const int *rowind, long long nnz, int nrows, int *flags) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
        const int r = rowind[i];
        if (r < 0 || r >= nrows) flags[1] = 1;
        if (i > 0 && rowind[i - 1] > r) flags[0] = 1;
    }
}
__global__ void coo_rowptr_kernel(const int *rowind, long long nnz, int nrows, int *rowptr) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= nnz; i += (long long)gridDim.x * blockDim.x) {
        const int lo = i == 0 ? 0 : rowind[i - 1] + 1;
        const int hi = i == nnz ? nrows : rowind[i];
        for (int r = lo; r <= hi; ++r) rowptr[r] = (int)i;
    }
}
// spins until every flag has reached `epoch` (the consumer side of the in-kernel arrival flags)
__global__ void wait_flags_kernel(const int *flags, int n, int epoch) {
    if ((int)threadIdx.x < n) {
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
            if (v - epoch < 0) __nanosleep(64);
        } while (v - epoch < 0);
    }
}

static cudaError_t dispatch_all_ones(int dtype, const void *val, long long n, int *flag) {
    switch (dtype) {
        case PYGIM_INT8: return check_all_ones_i8(val, n, flag, nullptr);
        case PYGIM_INT16: return check_all_ones_i16(val, n, flag, nullptr);
        case PYGIM_INT32: return check_all_ones_i32(val, n, flag, nullptr);
        case PYGIM_INT64: return check_all_ones_i64(val, n, flag, nullptr);
        case PYGIM_FLT32: return check_all_ones_f32(val, n, flag, nullptr);
        default: return check_all_ones_f64(val, n, flag, nullptr);
    }
}

template <typename L> static cudaError_t dispatch_csr(int dtype, const L &l, int64_t *n) {
    switch (dtype) {
        case PYGIM_INT8: return launch_csr_i8(l, n);
        case PYGIM_INT16: return launch_csr_i16(l, n);
        case PYGIM_INT32: return launch_csr_i32(l, n);
        case PYGIM_INT64: return launch_csr_i64(l, n);
        case PYGIM_FLT32: return launch_csr_f32(l, n);
        default: return launch_csr_f64(l, n);
    }
}
template <typename L> static cudaError_t dispatch_coo(int dtype, const L &l, int64_t *n) {
    switch (dtype) {
        case PYGIM_INT8: return launch_coo_i8(l, n);
        case PYGIM_INT16: return launch_coo_i16(l, n);
        case PYGIM_INT32: return launch_coo_i32(l, n);
        case PYGIM_INT64: return launch_coo_i64(l, n);
        case PYGIM_FLT32: return launch_coo_f32(l, n);
        default: return launch_coo_f64(l, n);
    }
}

// Access-policy window over the dense tile a launch gathers from: its lines are kept as "persisting" in the L2
// carve-out while the A stream (read once, evict-first) and the C stores pass through.  `touched` = bytes of the
// window the kernel really reads (a column tile touches only its own columns of every row).
static void set_l2_window(cudaStream_t stream, const void *base, size_t span_bytes, size_t touched_bytes) {
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof attr);
    if (base && span_bytes && g_ctx.persisting_set > 0 && g_ctx.max_window > 0) {
        attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
        attr.accessPolicyWindow.num_bytes = std::min<size_t>(span_bytes, (size_t)g_ctx.max_window);
        const double ratio = touched_bytes ? (double)g_ctx.persisting_set * 0.9 / (double)touched_bytes : 1.0;
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }   // else: num_bytes == 0 clears the window
    (void)cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    (void)cudaGetLastError();
}

// C[:, col0:col0+w] (+)= A_i * B_tile[rows_i, :]  for one (sparse part, dense tile[, row range of the plan]).
// `epi` (optional) already points at this tile's columns; C is then float32 when it de-quantises / adds a residual.
static int run_tile(Group *g, SparsePart &p, const char *B, long long ldb, char *C, long long ldc, long long width,
                    bool accumulate, cudaStream_t stream, const EpilogueLaunch *epi = nullptr, CsrPlan *plan = nullptr) {
    const size_t s = dtype_size(g->dtype);
    if (ldb < 0 || (unsigned long long)ldb * s >= (1ull << 32))
        return fail(PYGIM_ERR_INVALID, "row stride of the dense operand must be below 4 GiB");
    // Access-policy window (persisting L2 lines) over the tile the launch gathers from.  It halves the DRAM traffic of
    // a long launch (Reddit-shape, 64-column tile of H = 128: 1.91 -> 0.88 GB) without changing its time (sweep 9407 vs
    // 9404 GFLOP/s), but the lines it pins outlive the launch: when the next launch gathers from ANOTHER operand -
    // every sweep, every layer - they crowd the carve-out, and a short launch cannot amortise that.  Measured on a
    // 1/8 Reddit-shape row shard (what each GPU runs at N = 8), hidden sweep as one CUDA graph: 1166 us with the
    // window, 913 us without.  So: automatic = on only when the tile fits the carve-out AND every feature row is
    // gathered at least 200 times by the launch (whole Reddit-shape: 492; a 1/4 shard: 123); "l2_persist" forces it.
    const size_t touched = (size_t)p.ncols * (size_t)width * s;
    const bool persist = g->opt_l2_persist > 0 ||
                         (g->opt_l2_persist < 0 && touched >= (size_t)(4u << 20) &&
                          (double)touched <= 0.9 * (double)g_ctx.persisting_set && p.nnz >= 200 * std::max<long long>(p.ncols, 1));
    if (persist) set_l2_window(stream, B, (size_t)p.ncols * (size_t)ldb * s, touched);
    struct WindowGuard {      // the stream belongs to the caller: never leave our policy window behind
        cudaStream_t st; bool on;
        ~WindowGuard() { if (on) set_l2_window(st, nullptr, 0, 0); }
    } guard{stream, persist};
    cudaError_t err;
    const bool float_out = epi && (epi->scale || epi->residual);
    if (g->format == PYGIM_CSR || g->csr_view) {
        CsrPlan &pl = plan ? *plan : p.full;
        const long long ldp = (width * (long long)s + 15) / 16 * 16 / (long long)s;
        int max_g = g->opt_max_g > 0 ? (int)g->opt_max_g : 32;
        if (p.hot_k > 0) max_g = std::min(max_g, 8);      // hot/cold: 128-byte column chunks (one tile row = one line)
        const bool vec = csr_can_vectorize(s, B, C, float_out ? sizeof(float) : s, width, ldb, ldc, ldp);
        const int chunks = csr_col_chunks(s, width, max_g, vec);
        const size_t n_counters = 4 + (size_t)pl.n_super * chunks + (size_t)pl.n_long * chunks;
        Scratch scratch;
        Scratch *sc = &scratch;
        int rc = get_scratch(g, stream, n_counters, (size_t)pl.n_slots * (size_t)ldp * s, false, sc);
        if (rc) return rc;
        CsrLaunch l;
        l.rowptr = p.csr_rowptr() + pl.row_begin;
        l.colind = p.colind;
        l.val = p.values;
        l.B = B;
        l.C = C + (size_t)pl.row_begin * (size_t)ldc * (float_out ? sizeof(float) : s);
        l.partial = sc->partial;
        l.segs = pl.d_segs;
        l.supers = pl.d_supers;
        l.long_rows = pl.d_long_rows;
        l.long_seg_ptr = pl.d_long_seg_ptr;
        l.warps_out = reinterpret_cast<unsigned int *>(sc->counters);
        l.super_cnt = sc->counters + 4;
        l.seg_count = sc->counters + 4 + (size_t)pl.n_super * chunks;
        l.n_super = pl.n_super;
        l.n_items = pl.n_items;
        l.n_seg = pl.n_seg;
        l.n_long = pl.n_long;
        l.nrows = (int)(pl.row_end - pl.row_begin);
        l.seg_len = p.seg_len;
        l.nnz_total = p.nnz;
        {   // the family is part of the plan (kernel_family); a plan without the tiny-row split runs family 4 as 3
            const int family = kernel_family(*g, p);
            l.short_rows = pl.tiny_split ? 4 : (family == 4 ? 3 : family);
        }
        l.max_g = max_g;
        l.cta_threads = g->opt_cta_threads > 0 ? (int)g->opt_cta_threads : 256;
        if (p.hot_k > 0) {
            if (plan) return fail(PYGIM_ERR_INVALID, "hot/cold plans have no row-range sub-plans");
            l.hot_cols = p.d_hot_cols;
            l.hot_cnt = p.d_hot_cnt;
            l.hot_k = p.hot_k;
            l.n_seg_super = pl.n_seg_super;
        }
        l.ncols = width;
        l.ldb = ldb;
        l.ldc = ldc;
        l.ldp = ldp;
        l.accumulate = accumulate ? 1 : 0;
        l.unit_values = (p.unit_values && g->opt_unit_values != 0) ? 1 : 0;
        if (epi) l.epi = *epi;
        if (g->d_row_map) l.epi.row_map = g->d_row_map + pl.row_begin;
        if (l.epi.peer_mask) l.epi.peer_mask += pl.row_begin;
        l.sm_count = g_ctx.sm_count;
        l.stream = stream;
        err = dispatch_csr(g->dtype, l, &g->last_launches);
    } else {
        if (float_out || (epi && epi->n_peers > 0) || g->d_row_map)
            return fail(PYGIM_ERR_INVALID, "fused epilogues / row maps need a CSR plan or a row-major sorted COO plan");
        Scratch scratch;
        Scratch *sc = &scratch;
        int rc = get_scratch(g, stream, 0, 0, true, sc);
        if (rc) return rc;
        CooLaunch l;
        l.rowind = p.rowidx;
        l.colind = p.colind;
        l.val = p.values;
        l.B = B;
        l.C = C;
        l.nnz = p.nnz;
        l.nrows = p.nrows;
        l.ncols = width;
        l.ldb = ldb;
        l.ldc = ldc;
        l.chunk_nnz = (int)g->opt_chunk_nnz;
        l.unit_values = (p.unit_values && g->opt_unit_values != 0) ? 1 : 0;
        l.accumulate = accumulate ? 1 : 0;
        l.all_atomic = p.coo_sorted ? 0 : 1;
        l.n_warp_slots = g_ctx.sm_count * (g_ctx.max_threads_per_sm / 32);
        l.sm_count = g_ctx.sm_count;
        l.ticket = sc->coo_ticket;
        l.stream = stream;
        err = dispatch_coo(g->dtype, l, &g->last_launches);
    }
    if (err != cudaSuccess) return fail(PYGIM_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(err));
    return PYGIM_OK;
}

// The (sparse part x dense part) loop of spmm_pim_csr / spmm_host_*_group (ops.hpp:42-62):
// dense part j of width h_j lands at column offset sum_{k<j} h_k; sparse part 0 overwrites, parts >= 1 add.
// `epi` describes the whole result matrix; its pointers are advanced to each column tile here.
static int run_group_device(Group *g, int n_ds, const void *const *B_parts, const long long *ldb, void *C,
                            long long ldc, cudaStream_t stream, const EpilogueLaunch *epi = nullptr, int chunk = -1) {
    if (n_ds != (int)g->dense_cols.size())
        return fail(PYGIM_ERR_INVALID, "expected %d dense parts, got %d", (int)g->dense_cols.size(), n_ds);
    const size_t s = dtype_size(g->dtype);
    const bool float_out = epi && (epi->scale || epi->residual);
    const size_t so = float_out ? sizeof(float) : s;
    if (epi && (float_out || epi->n_peers > 0) && g->parts.size() != 1)
        // partial products of sparse parts >= 1 would need a (remote) read-modify-write of a de-quantised value
        return fail(PYGIM_ERR_INVALID, "fused epilogues (de-quantise, residual, all-gather) need sp_parts == 1");
    if (chunk <= 0) g->last_launches = 0;
    long long brow = 0;
    for (size_t i = 0; i < g->parts.size(); ++i) {
        long long ccol = 0;
        for (int j = 0; j < n_ds; ++j) {
            const long long w = g->dense_cols[j];
            const char *B = static_cast<const char *>(B_parts[j]) + (size_t)brow * (size_t)ldb[j] * s;
            char *Ct = C ? static_cast<char *>(C) + (size_t)ccol * so : nullptr;
            EpilogueLaunch e;
            if (epi) {
                e = *epi;
                for (int q = 0; q < e.n_peers; ++q) e.peers[q] = static_cast<char *>(e.peers[q]) + (size_t)ccol * so;
                if (e.mc) e.mc = static_cast<char *>(e.mc) + (size_t)ccol * so;
                if (e.residual) e.residual += ccol;
                if (e.n_peers > 0) Ct = static_cast<char *>(e.peers[0]);
                // only the LAST launch of the call announces the rows (flags are per call, not per tile)
                if (!(i + 1 == g->parts.size() && j + 1 == n_ds))
                    for (int q = 0; q < 8; ++q) e.flags[q] = nullptr;
            }
            int rc = run_tile(g, g->parts[i], B, ldb[j], Ct, ldc, w, i > 0, stream, epi ? &e : nullptr,
                              chunk >= 0 ? &g->parts[i].chunks[chunk] : nullptr);
            if (rc) return rc;
            ccol += w;
        }
        brow += g->parts[i].ncols;
    }
    return PYGIM_OK;
}

}  // namespace pygim

using namespace pygim;

int pygim_fail_invalid(const char *msg) { return fail(PYGIM_ERR_INVALID, "%s", msg); }

// =============================================================================== C ABI
extern "C" {

PYGIM_API const char *pygim_last_error(void) { return g_err.c_str(); }
PYGIM_API int pygim_abi_version(void) { return 2; }

PYGIM_API int pygim_dpu_init_ranks(int64_t nr_ranks, int64_t groups_per_rank, int device, int32_t *units_per_rank_out) {
    if (nr_ranks <= 0) return fail(PYGIM_ERR_INVALID, "nr_ranks must be positive, got %lld", (long long)nr_ranks);
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    int rc = context_init(device);
    if (rc) return rc;
    g_ctx.nr_ranks = nr_ranks;
    g_ctx.groups_per_rank = groups_per_rank > 0 ? groups_per_rank : 1;
    g_ctx.nr_dpus = 0;
    if (units_per_rank_out) {
        const int per = std::max<long long>(1, g_ctx.sm_count / nr_ranks);
        for (int64_t i = 0; i < nr_ranks; ++i) units_per_rank_out[i] = per;
    }
    return PYGIM_OK;
}

PYGIM_API int pygim_dpu_init_dpus(int64_t nr_dpus, int device) {
    if (nr_dpus <= 0) return fail(PYGIM_ERR_INVALID, "nr_dpus must be positive, got %lld", (long long)nr_dpus);
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    int rc = context_init(device);
    if (rc) return rc;
    g_ctx.nr_dpus = nr_dpus;
    g_ctx.nr_ranks = 1;
    return PYGIM_OK;
}

PYGIM_API int pygim_dpu_release(void) {
    std::lock_guard<std::mutex> lock(g_ctx.mu);
    if (g_ctx.initialised && g_ctx.persisting_set > 0) {   // hand the L2 back: no stale persisting lines
        (void)cudaCtxResetPersistingL2Cache();
        (void)cudaGetLastError();
    }
    g_ctx.initialised = false;
    g_ctx.nr_ranks = g_ctx.nr_dpus = 0;
    return PYGIM_OK;
}

PYGIM_API int pygim_device_info(int *device, int *sm_count, int64_t *l2_bytes, int64_t *persisting_l2_max_bytes,
                      int64_t *hbm_bytes, int *cc_major, int *cc_minor) {
    if (!g_ctx.initialised) return fail(PYGIM_ERR_NOT_INIT, "call pygim_dpu_init_ranks / pygim_dpu_init_dpus first");
    if (device) *device = g_ctx.device;
    if (sm_count) *sm_count = g_ctx.sm_count;
    if (l2_bytes) *l2_bytes = g_ctx.l2_bytes;
    if (persisting_l2_max_bytes) *persisting_l2_max_bytes = g_ctx.persisting_max;
    if (hbm_bytes) *hbm_bytes = g_ctx.hbm_bytes;
    if (cc_major) *cc_major = g_ctx.cc_major;
    if (cc_minor) *cc_minor = g_ctx.cc_minor;
    return PYGIM_OK;
}

PYGIM_API int pygim_spmm_to_device_group(int format, int dtype, int n_sp, const int32_t *const *rowidx,
                               const int32_t *const *colind, const void *const *values, const int64_t *nrows,
                               const int64_t *ncols, const int64_t *nnz, int n_ds, const int64_t *dense_cols,
                               int64_t h_size, int mem, pygim_handle_t *out_handle) {
    if (!out_handle) return fail(PYGIM_ERR_INVALID, "out_handle is NULL");
    *out_handle = 0;
    if (!g_ctx.initialised) return fail(PYGIM_ERR_NOT_INIT, "call pygim_dpu_init_ranks / pygim_dpu_init_dpus first");
    if (format != PYGIM_CSR && format != PYGIM_COO) return fail(PYGIM_ERR_INVALID, "unknown format %d", format);
    const size_t s = dtype_size(dtype);
    if (s == 0) return fail(PYGIM_ERR_INVALID, "unknown dtype %d", dtype);
    if (n_sp <= 0 || n_ds <= 0) return fail(PYGIM_ERR_INVALID, "need at least one sparse and one dense part");
    if (mem != PYGIM_MEM_HOST && mem != PYGIM_MEM_DEVICE) return fail(PYGIM_ERR_INVALID, "unknown mem kind %d", mem);
    long long hsum = 0;
    for (int j = 0; j < n_ds; ++j) {
        if (dense_cols[j] < 0) return fail(PYGIM_ERR_INVALID, "negative dense part width");
        hsum += dense_cols[j];
    }
    if (hsum != h_size)   // the reference assert()s this at run time (pytorch_api.cpp:266)
        return fail(PYGIM_ERR_INVALID, "dense part widths sum to %lld, h_size is %lld", hsum, (long long)h_size);
    for (int i = 0; i < n_sp; ++i) {
        if (nrows[i] != nrows[0]) return fail(PYGIM_ERR_INVALID, "sparse parts must share the row count (col_split)");
        if (nrows[i] < 0 || ncols[i] < 0 || nnz[i] < 0 || nnz[i] > 0x7fffffffLL || nrows[i] >= 0x7fffffffLL)
            return fail(PYGIM_ERR_INVALID, "sparse part %d: sizes out of the int32 index range", i);
    }
    CUDA_TRY(cudaSetDevice(g_ctx.device));

    Group *g = new Group;
    g->format = format;
    g->dtype = dtype;
    g->device = g_ctx.device;
    g->h_size = h_size;
    g->total_rows = nrows[0];
    g->dense_cols.assign(dense_cols, dense_cols + n_ds);
    g->parts.resize(n_sp);
    auto bail = [&](int rc) { destroy_group(g); return rc; };

    for (int i = 0; i < n_sp; ++i) {
        SparsePart &p = g->parts[i];
        p.nrows = nrows[i];
        p.ncols = ncols[i];
        p.nnz = nnz[i];
        g->total_cols += ncols[i];
        const size_t ridx_n = format == PYGIM_CSR ? (size_t)nrows[i] + 1 : (size_t)nnz[i];
        if (mem == PYGIM_MEM_HOST) {
            int *d_r = nullptr, *d_c = nullptr;
            void *d_v = nullptr;
            p.owned = true;
            cudaError_t e;
            if ((e = cudaMalloc(&d_r, std::max<size_t>(ridx_n, 1) * 4)) != cudaSuccess ||
                (e = cudaMalloc(&d_c, std::max<size_t>((size_t)nnz[i], 1) * 4)) != cudaSuccess ||
                (e = cudaMalloc(&d_v, std::max<size_t>((size_t)nnz[i], 1) * s)) != cudaSuccess) {
                p.rowidx = d_r; p.colind = d_c; p.values = d_v;
                return bail(fail(PYGIM_ERR_CUDA, "cudaMalloc of sparse part %d failed: %s", i, cudaGetErrorString(e)));
            }
            p.rowidx = d_r; p.colind = d_c; p.values = d_v;
            if ((e = cudaMemcpy(d_r, rowidx[i], ridx_n * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
                (e = cudaMemcpy(d_c, colind[i], (size_t)nnz[i] * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
                (e = cudaMemcpy(d_v, values[i], (size_t)nnz[i] * s, cudaMemcpyHostToDevice)) != cudaSuccess)
                return bail(fail(PYGIM_ERR_CUDA, "upload of sparse part %d failed: %s", i, cudaGetErrorString(e)));
        } else {
            p.owned = false;
            p.rowidx = rowidx[i];
            p.colind = colind[i];
            p.values = values[i];
        }
        {   // unit-value detection (device-side scan of the values as they will be read)
            int *d_flag = nullptr, h_flag = 1;
            cudaError_t e = cudaMalloc(&d_flag, sizeof(int));
            if (e == cudaSuccess) e = cudaMemcpy(d_flag, &h_flag, sizeof(int), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = dispatch_all_ones(dtype, p.values, p.nnz, d_flag);
            if (e == cudaSuccess) e = cudaMemcpy(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost);
            if (d_flag) cudaFree(d_flag);
            if (e != cudaSuccess)
                return bail(fail(PYGIM_ERR_CUDA, "unit-value scan of sparse part %d failed: %s", i, cudaGetErrorString(e)));
            p.unit_values = (h_flag == 1) && p.nnz > 0;
        }
        if (format == PYGIM_CSR) {
            p.h_rowptr.resize((size_t)nrows[i] + 1);
            if (mem == PYGIM_MEM_HOST) {
                std::memcpy(p.h_rowptr.data(), rowidx[i], ((size_t)nrows[i] + 1) * 4);
            } else {
                cudaError_t e = cudaMemcpy(p.h_rowptr.data(), rowidx[i], ((size_t)nrows[i] + 1) * 4, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess)
                    return bail(fail(PYGIM_ERR_CUDA, "read-back of rowptr failed: %s", cudaGetErrorString(e)));
            }
            if ((long long)(unsigned)p.h_rowptr[(size_t)nrows[i]] != nnz[i] || p.h_rowptr[0] != 0)
                return bail(fail(PYGIM_ERR_INVALID, "sparse part %d: rowptr does not span [0, nnz]", i));
        } else {
            // COO: is the stream row-major sorted (what spmm.py:40-42 `.coalesce()` yields)?  Row ids in range?
            int *d_flags = nullptr, h_flags[2] = {0, 0};
            cudaError_t e = cudaMalloc(&d_flags, 2 * sizeof(int));
            if (e == cudaSuccess) e = cudaMemset(d_flags, 0, 2 * sizeof(int));
            if (e == cudaSuccess && p.nnz > 0) {
                coo_check_sorted_kernel<<<1184, 256>>>(p.rowidx, p.nnz, (int)p.nrows, d_flags);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaMemcpy(h_flags, d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost);
            if (d_flags) cudaFree(d_flags);
            if (e != cudaSuccess)
                return bail(fail(PYGIM_ERR_CUDA, "sortedness scan of sparse part %d failed: %s", i, cudaGetErrorString(e)));
            if (h_flags[1]) return bail(fail(PYGIM_ERR_INVALID, "sparse part %d: COO row index out of range", i));
            p.coo_sorted = h_flags[0] == 0;
            if (p.coo_sorted) {
                // derive the row pointer once: a sorted COO stream then runs through the CSR kernels (rowind is
                // not read again), an unsorted one keeps the COO kernel with every flush atomic
                e = cudaMalloc(reinterpret_cast<void **>(&p.d_rowptr), ((size_t)p.nrows + 1) * sizeof(int));
                if (e == cudaSuccess) {
                    coo_rowptr_kernel<<<1184, 256>>>(p.rowidx, p.nnz, (int)p.nrows, p.d_rowptr);
                    e = cudaGetLastError();
                }
                p.h_rowptr.resize((size_t)p.nrows + 1);
                if (e == cudaSuccess)
                    e = cudaMemcpy(p.h_rowptr.data(), p.d_rowptr, ((size_t)p.nrows + 1) * 4, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess)
                    return bail(fail(PYGIM_ERR_CUDA, "row pointer of sparse part %d failed: %s", i, cudaGetErrorString(e)));
            }
        }
    }
    g->csr_view = format == PYGIM_COO;
    for (auto &p : g->parts) g->csr_view = g->csr_view && p.coo_sorted;
    if (format == PYGIM_COO && !g->csr_view)
        for (auto &p : g->parts) { p.h_rowptr.clear(); }
    {
        int rc = replan(g);
        if (rc) return bail(rc);
    }
    for (auto &e : g->ev) {
        cudaError_t ce = cudaEventCreate(&e);
        if (ce != cudaSuccess) return bail(fail(PYGIM_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(ce)));
    }
    {
        std::lock_guard<std::mutex> lock(g_reg_mu);
        *out_handle = g_next_handle++;
        g_registry[*out_handle] = g;
    }
    return PYGIM_OK;
}

PYGIM_API int pygim_spmm_free_group(pygim_handle_t handle) {
    Group *g = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_reg_mu);
        auto it = g_registry.find(handle);
        if (it == g_registry.end())
            return fail(PYGIM_ERR_INVALID, "unknown or freed plan handle %llu", (unsigned long long)handle);
        g = it->second;
        g_registry.erase(it);
    }
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    destroy_group(g);
    return PYGIM_OK;
}

PYGIM_API int pygim_plan_set_option(pygim_handle_t handle, const char *key, int64_t value) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!key) return fail(PYGIM_ERR_INVALID, "null key");
    bool rebuild = false;
    if (!std::strcmp(key, "seg_len")) {
        g->opt_seg_len = value; rebuild = true;
    } else if (!std::strcmp(key, "item_nnz")) {
        g->opt_item_nnz = value; rebuild = true;
    } else if (!std::strcmp(key, "super_nnz")) {
        g->opt_super_nnz = value; rebuild = true;
    } else if (!std::strcmp(key, "rows_per_ticket")) {
        g->opt_rows_per_ticket = value; rebuild = true;
    } else if (!std::strcmp(key, "max_g")) {
        if (value > 0 && (value > 32 || (value & (value - 1)))) return fail(PYGIM_ERR_INVALID, "max_g must be a power of two <= 32");
        g->opt_max_g = value;
    } else if (!std::strcmp(key, "cta_threads")) {
        if (value > 1024) return fail(PYGIM_ERR_INVALID, "cta_threads must be <= 1024");
        g->opt_cta_threads = value;
    } else if (!std::strcmp(key, "l2_persist")) {
        g->opt_l2_persist = value;
    } else if (!std::strcmp(key, "chunk_nnz")) {
        g->opt_chunk_nnz = value;
    } else if (!std::strcmp(key, "unit_values")) {
        g->opt_unit_values = value;
    } else if (!std::strcmp(key, "short_rows")) {
        if (g->opt_short_rows != value && !(g->parts.empty() || g->parts[0].hot_k > 0)) rebuild = true;   // the family shapes the plan
        g->opt_short_rows = value;
    } else if (!std::strcmp(key, "host_chunks")) {
        g->opt_host_chunks = value;
    } else if (!std::strcmp(key, "host_tile_bytes")) {
        if (value > 0 && value % 128) return fail(PYGIM_ERR_INVALID, "host_tile_bytes must be a multiple of 128");
        g->opt_host_tile_bytes = value;
    } else if (!std::strcmp(key, "coo_native")) {
        if (g->format != PYGIM_COO) return fail(PYGIM_ERR_INVALID, "coo_native applies to COO plans");
        g->opt_coo_native = value;
        bool sorted = true;
        for (auto &p : g->parts) sorted = sorted && p.coo_sorted && p.d_rowptr;
        g->csr_view = sorted && value <= 0;
    } else {
        return fail(PYGIM_ERR_INVALID, "unknown option '%s'", key);
    }
    if (rebuild && !g->parts.empty() && g->parts[0].hot_k > 0)
        return fail(PYGIM_ERR_INVALID, "option '%s' must be set before pygim_plan_set_hot_tiles", key);
    if (rebuild && (g->format == PYGIM_CSR || g->csr_view)) {
        CUDA_TRY(cudaSetDevice(g->device));
        CUDA_TRY(cudaDeviceSynchronize());
        return replan(g);
    }
    return PYGIM_OK;
}

PYGIM_API int pygim_plan_stats(pygim_handle_t handle, int part, int64_t *out8) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!out8) return fail(PYGIM_ERR_INVALID, "null output");
    if (part < 0 || part >= (int)g->parts.size()) return fail(PYGIM_ERR_INVALID, "part %d out of range", part);
    const SparsePart &p = g->parts[part];
    out8[0] = p.nrows;
    out8[1] = p.ncols;
    out8[2] = p.nnz;
    out8[3] = p.max_row_nnz;
    out8[4] = p.full.n_long;
    out8[5] = p.full.n_seg;
    out8[6] = p.seg_len;
    out8[7] = p.empty_rows;
    return PYGIM_OK;
}

PYGIM_API int pygim_plan_layout(pygim_handle_t handle, int part, int64_t *out6) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!out6) return fail(PYGIM_ERR_INVALID, "null output");
    if (part < 0 || part >= (int)g->parts.size()) return fail(PYGIM_ERR_INVALID, "part %d out of range", part);
    const SparsePart &p = g->parts[part];
    out6[0] = p.full.n_items;
    out6[1] = p.full.n_super;
    out6[2] = g->csr_view ? 1 : 0;
    out6[3] = p.coo_sorted ? 1 : 0;
    out6[4] = g->d_row_map ? 1 : 0;
    out6[5] = p.unit_values ? 1 : 0;
    return PYGIM_OK;
}

PYGIM_API int pygim_plan_set_row_map(pygim_handle_t handle, const int32_t *row_map, int64_t n, int mem) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    CUDA_TRY(cudaSetDevice(g->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (g->d_row_map) { cudaFree(g->d_row_map); g->d_row_map = nullptr; }
    if (!row_map) return g->parts[0].hot_k == 0 ? replan(g) : PYGIM_OK;
    if (n != g->total_rows) return fail(PYGIM_ERR_INVALID, "row map has %lld entries, the plan has %lld rows", (long long)n, g->total_rows);
    if (g->format == PYGIM_COO && !g->csr_view) return fail(PYGIM_ERR_INVALID, "row maps need a CSR plan or a row-major sorted COO plan");
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&g->d_row_map), std::max<size_t>((size_t)n, 1) * sizeof(int)));
    CUDA_TRY(cudaMemcpy(g->d_row_map, row_map, (size_t)n * sizeof(int),
                        mem == PYGIM_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
    if (g->parts[0].hot_k == 0) return replan(g);      // a reordered plan schedules SM-affine supertickets
    return PYGIM_OK;
}

PYGIM_API int pygim_plan_set_hot_tiles(pygim_handle_t handle, int64_t n_super, const int32_t *super_rows, int hot_k,
                                       const int32_t *hot_cols, const int32_t *hot_cnt, int mem) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (g->parts.size() != 1) return fail(PYGIM_ERR_INVALID, "hot/cold plans need sp_parts == 1");
    if (g->format == PYGIM_COO && !g->csr_view) return fail(PYGIM_ERR_INVALID, "hot/cold plans need a CSR plan or a sorted COO plan");
    if (!super_rows || !hot_cols || !hot_cnt || n_super <= 0 || hot_k <= 0) return fail(PYGIM_ERR_INVALID, "bad hot-tile arguments");
    SparsePart &p = g->parts[0];
    CUDA_TRY(cudaSetDevice(g->device));
    CUDA_TRY(cudaDeviceSynchronize());
    std::vector<int> rows((size_t)n_super + 1);
    const cudaMemcpyKind to_host = mem == PYGIM_MEM_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost;
    const cudaMemcpyKind to_dev = mem == PYGIM_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CUDA_TRY(cudaMemcpy(rows.data(), super_rows, rows.size() * sizeof(int), to_host));
    if (rows[0] != 0 || rows[(size_t)n_super] != (int)p.nrows) return fail(PYGIM_ERR_INVALID, "superticket rows must span [0, nrows]");
    for (int64_t k = 0; k < n_super; ++k)
        if (rows[(size_t)k + 1] <= rows[(size_t)k]) return fail(PYGIM_ERR_INVALID, "superticket %lld is empty", (long long)k);
    if ((size_t)hot_k * 8 * 16 > (size_t)200 * 1024) return fail(PYGIM_ERR_INVALID, "hot_k %d does not fit shared memory", hot_k);
    if (p.d_hot_cols) { cudaFree(p.d_hot_cols); p.d_hot_cols = nullptr; }
    if (p.d_hot_cnt) { cudaFree(p.d_hot_cnt); p.d_hot_cnt = nullptr; }
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&p.d_hot_cols), (size_t)n_super * hot_k * sizeof(int)));
    CUDA_TRY(cudaMemcpy(p.d_hot_cols, hot_cols, (size_t)n_super * hot_k * sizeof(int), to_dev));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&p.d_hot_cnt), std::max<size_t>((size_t)p.nrows, 1) * sizeof(int)));
    CUDA_TRY(cudaMemcpy(p.d_hot_cnt, hot_cnt, (size_t)p.nrows * sizeof(int), to_dev));
    p.hot_super_rows = rows;
    p.hot_k = 0;                       // replan with the fixed boundaries, then switch the mode on
    int rc = replan(g);
    if (rc) return rc;
    p.hot_k = hot_k;
    return PYGIM_OK;
}

// B as one matrix: the dense parts are its column tiles
static void column_tiles(const Group *g, const void *B, int64_t ldb, std::vector<const void *> &parts, std::vector<long long> &lds) {
    const size_t s = dtype_size(g->dtype);
    parts.resize(g->dense_cols.size());
    lds.assign(g->dense_cols.size(), ldb);
    long long col = 0;
    for (size_t j = 0; j < g->dense_cols.size(); ++j) {
        parts[j] = static_cast<const char *>(B) + (size_t)col * s;
        col += g->dense_cols[j];
    }
}

PYGIM_API int pygim_spmm_run_group_device(pygim_handle_t handle, int n_ds, const void *const *B_parts, const int64_t *ldb,
                                void *C, int64_t ldc, void *stream) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!B_parts || !ldb || !C) return fail(PYGIM_ERR_INVALID, "null buffer");
    std::vector<long long> l(ldb, ldb + n_ds);
    return run_group_device(g, n_ds, B_parts, l.data(), C, ldc, static_cast<cudaStream_t>(stream));
}

PYGIM_API int pygim_spmm_device(pygim_handle_t handle, const void *B, int64_t ldb, void *C, int64_t ldc, void *stream) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!B || !C) return fail(PYGIM_ERR_INVALID, "null buffer");
    std::vector<const void *> parts;
    std::vector<long long> lds;
    column_tiles(g, B, ldb, parts, lds);
    return run_group_device(g, (int)parts.size(), parts.data(), lds.data(), C, ldc, static_cast<cudaStream_t>(stream));
}

PYGIM_API int pygim_spmm_device_ex(pygim_handle_t handle, const void *B, int64_t ldb, void *C, int64_t ldc,
                                   const pygim_epilogue_t *epi, void *stream) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!epi) return pygim_spmm_device(handle, B, ldb, C, ldc, stream);
    if (!B) return fail(PYGIM_ERR_INVALID, "null buffer");
    if (epi->n_peers < 0 || epi->n_peers > 8) return fail(PYGIM_ERR_INVALID, "n_peers must be 0..8, got %d", epi->n_peers);
    if (epi->n_peers == 0 && !C) return fail(PYGIM_ERR_INVALID, "null result buffer");
    if (epi->n_peers > 0 && !epi->C_peers) return fail(PYGIM_ERR_INVALID, "n_peers > 0 needs C_peers");
    if (epi->residual && g->dtype != PYGIM_FLT32 && !epi->scale)
        return fail(PYGIM_ERR_INVALID, "a residual needs a float32 result (FLT32 plan, or scale set)");
    const bool float_out = epi->scale || epi->residual;
    const size_t so = float_out ? sizeof(float) : dtype_size(g->dtype);
    EpilogueLaunch e;
    e.scale = epi->scale;
    e.residual = epi->residual;
    e.ld_res = epi->ld_residual;
    e.coeff = epi->residual_coeff;
    e.n_peers = epi->n_peers;
    const size_t row_bytes = (size_t)epi->row_offset * (size_t)ldc * so;
    for (int q = 0; q < epi->n_peers; ++q) {
        if (!epi->C_peers[q]) return fail(PYGIM_ERR_INVALID, "peer %d has a null buffer", q);
        e.peers[q] = static_cast<char *>(epi->C_peers[q]) + row_bytes;
        e.flags[q] = epi->flag_peers ? epi->flag_peers[q] : nullptr;
    }
    e.mc = epi->C_multicast ? static_cast<char *>(epi->C_multicast) + row_bytes : nullptr;
    e.peer_mask = epi->row_peer_mask;
    e.my_rank = epi->my_rank;
    e.epoch = epi->epoch;
    if (e.residual) e.residual += (size_t)epi->row_offset * (size_t)e.ld_res;
    std::vector<const void *> parts;
    std::vector<long long> lds;
    column_tiles(g, B, ldb, parts, lds);
    return run_group_device(g, (int)parts.size(), parts.data(), lds.data(), C, ldc, static_cast<cudaStream_t>(stream), &e);
}

PYGIM_API int pygim_spmm_device_peers(pygim_handle_t handle, const void *B, int64_t ldb, void *const *C_peers,
                                      int n_peers, void *C_multicast, int64_t ldc, int64_t row_offset, void *stream) {
    if (n_peers < 1 || n_peers > 8) return fail(PYGIM_ERR_INVALID, "n_peers must be 1..8, got %d", n_peers);
    if (!C_peers) return fail(PYGIM_ERR_INVALID, "null buffer");
    pygim_epilogue_t epi;
    std::memset(&epi, 0, sizeof epi);
    epi.C_peers = C_peers;
    epi.n_peers = n_peers;
    epi.C_multicast = C_multicast;
    epi.row_offset = row_offset;
    return pygim_spmm_device_ex(handle, B, ldb, nullptr, ldc, &epi, stream);
}

PYGIM_API int pygim_wait_flags(const int32_t *flags, int n, int32_t epoch, void *stream) {
    if (!flags || n < 1 || n > 32) return fail(PYGIM_ERR_INVALID, "bad flag arguments");
    wait_flags_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, n, epoch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PYGIM_ERR_CUDA, "wait_flags launch failed: %s", cudaGetErrorString(e));
    return PYGIM_OK;
}

// Host entry point.  Software pipeline over three streams so that PCIe and the kernels overlap:
//   * every dense part is cut into column tiles of 128 bytes per row (when it is at least 256 bytes wide); tile
//     t+1 is uploaded (copy-in stream) while tile t is computed, and the finished C tile t is downloaded
//     (copy-out stream) while tile t+1 is computed;
//   * the LAST tile (or the only one) is additionally cut into nnz-balanced row chunks (CSR) so that its download
//     overlaps its own kernels and only the last chunk's copy is exposed.
// Each tile is staged contiguously on the device ([rowsB x tile_width]), so it is L2 friendly for the gathers.
// Enqueue one host-operand SpMM as a three-stream software pipeline: column tile t+1 uploads on `sin` and the
// finished rows of tile t-1 download on `sout` while tile t computes on `st`.  Nothing is awaited: the caller joins
// `sout` into `st` and synchronises - once per call (pygim_spmm_run_group_host) or once per batch of calls
// (pygim_spmm_run_many_host, where the uploads of call k+1 overlap the kernels of call k as well).
// `chunk_last`: cut the launches of the LAST TWO tiles into row chunks whose downloads follow chunk by chunk, so
// only a fraction of one tile's download is exposed at the end (in a batch: the last call only - every other
// download already overlaps the kernels that follow it, and a chunked tile costs three extra launches).
static int host_enqueue(Group *g, int n_ds, const void *const *B_parts, const int64_t *ldb, void *C, int64_t ldc,
                        cudaStream_t st, cudaStream_t sin, cudaStream_t sout, bool first_of_batch, bool chunk_last) {
    if (n_ds != (int)g->dense_cols.size())
        return fail(PYGIM_ERR_INVALID, "expected %d dense parts, got %d", (int)g->dense_cols.size(), n_ds);
    CUDA_TRY(cudaSetDevice(g->device));
    const size_t s = dtype_size(g->dtype);
    const size_t rowsB = (size_t)g->total_cols, rowsC = (size_t)g->total_rows, H = (size_t)g->h_size;
    const bool pipelined = g->opt_host_chunks != 0 &&
                           (g->opt_host_chunks > 0 || rowsC * H * s >= (size_t)(8u << 20));

    // ---- column tiles.  Measured on the B200 box (tools/pcie_probe.py, profiles/r02_pcie_probe.txt): 2-D copies of
    // >= 256-byte row pieces run at the full PCIe rate (55 GB/s alone, 50 GB/s per direction with both busy); 128-byte
    // pieces reach 51 / 47 GB/s alone but only 35 GB/s per direction when uploads and downloads overlap - and in a
    // pipeline they always do (sweep step: 7.85 ms with 128-byte tiles, 7.47 ms with 256-byte tiles; thin tiles only
    // at the ends of the batch: 8.1 ms).  So tiles are 256 bytes wide (64 FLT32 columns: also the widest tile whose
    // gathers stay L2-resident on Reddit-shape); a remainder is one 128-byte tile.
    struct Tile { int part; size_t col0, width, dev_off; const char *host; size_t host_ld; };
    std::vector<Tile> tiles;
    size_t col = 0, dev_off = 0;
    const size_t wide = g->opt_host_tile_bytes > 0 ? (size_t)g->opt_host_tile_bytes : 256;
    const size_t thin = std::min<size_t>(wide, 128);
    for (int j = 0; j < n_ds; ++j) {
        const size_t w = (size_t)g->dense_cols[j];
        std::vector<size_t> widths;     // in bytes
        if (pipelined && w * s >= 2 * thin && (w * s) % thin == 0) {
            size_t left = w * s;
            while (left > 0) {
                const size_t t = std::min(left, (left % wide) ? thin : wide);
                widths.push_back(t);
                left -= t;
            }
        } else {
            widths.push_back(w * s);
        }
        size_t off = 0;
        for (size_t t = 0; t < widths.size(); ++t) {
            Tile tl;
            tl.part = j;
            tl.col0 = col + off / s;
            tl.width = widths[t] / s;
            tl.dev_off = dev_off;
            tl.host = static_cast<const char *>(B_parts[j]) + off;
            tl.host_ld = (size_t)ldb[j];
            if (tl.width) tiles.push_back(tl);
            dev_off += rowsB * widths[t];
            off += widths[t];
        }
        col += w;
    }
    const size_t needB = std::max<size_t>(dev_off, 16), needC = std::max<size_t>(rowsC * H * s, 16);
    if (needB > g->dB_bytes) {
        if (g->d_B) CUDA_TRY(cudaFree(g->d_B));
        g->d_B = nullptr; g->dB_bytes = 0;
        CUDA_TRY(cudaMalloc(&g->d_B, needB));
        g->dB_bytes = needB;
    }
    if (needC > g->dC_bytes) {
        if (g->d_C) CUDA_TRY(cudaFree(g->d_C));
        g->d_C = nullptr; g->dC_bytes = 0;
        CUDA_TRY(cudaMalloc(&g->d_C, needC));
        g->dC_bytes = needC;
    }

    // ---- row chunks of the last tile (CSR only: COO chunks are nnz ranges, not row ranges)
    int n_chunks = 1;
    if (chunk_last && pipelined && (g->format == PYGIM_CSR || g->csr_view) && !g->d_row_map && g->parts[0].hot_k == 0 &&
        g->parts[0].nnz >= 4096 && rowsC >= 64) {
        n_chunks = g->opt_host_chunks > 0 ? (int)g->opt_host_chunks : 4;
        if (g->parts[0].chunks.size() != (size_t)n_chunks) {
            std::vector<int64_t> split((size_t)n_chunks + 1);
            int rc = PYGIM_OK;
            if (g->opt_host_chunks > 0) {
                rc = pygim_partition_rows_by_nnz(g->parts[0].h_rowptr.data(), g->parts[0].nrows, n_chunks, split.data());
            } else {
                // automatic: shrinking chunks (40 / 30 / 20 / 10 % of the nonzeros) - only the LAST chunk's download is
                // exposed, so it should be the smallest, while four equal launches' worth of kernel efficiency is kept
                const std::vector<int> &rp = g->parts[0].h_rowptr;
                const long long rows = g->parts[0].nrows, total = (long long)(unsigned)rp[(size_t)rows];
                const double cum[5] = {0.0, 0.4, 0.7, 0.9, 1.0};
                split[0] = 0;
                for (int k = 1; k < 4; ++k) {
                    const long long target = (long long)(cum[k] * (double)total);
                    const long long r = std::lower_bound(rp.begin(), rp.begin() + rows + 1, target,
                                                         [](int v, long long t) { return (long long)(unsigned)v < t; }) - rp.begin();
                    split[(size_t)k] = std::min<long long>(rows, std::max<long long>(r, split[(size_t)k - 1]));
                }
                split[4] = rows;
            }
            if (rc) return rc;
            for (auto &p : g->parts) {
                for (auto &c : p.chunks) free_plan(c);
                p.chunks.assign((size_t)n_chunks, CsrPlan());
                for (int k = 0; k < n_chunks; ++k) {
                    rc = build_plan_range(*g, p, p.seg_len, split[k], split[k + 1], p.chunks[k]);
                    if (rc) return rc;
                }
            }
        }
    }
    while (g->chunk_done.size() < tiles.size() * (size_t)(n_chunks + 1) + 4) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g->chunk_done.push_back(e);
    }
    size_t next_event = 0;
    auto new_event = [&]() { return g->chunk_done[next_event++]; };

    g->last_launches = 0;
    CUDA_TRY(cudaEventRecord(g->ev[0], st));
    if (first_of_batch) {   // the copy streams start after everything already queued on the compute stream
        cudaEvent_t e = new_event();
        CUDA_TRY(cudaEventRecord(e, st));
        CUDA_TRY(cudaStreamWaitEvent(sin, e, 0));
        CUDA_TRY(cudaStreamWaitEvent(sout, e, 0));
    }
    std::vector<cudaEvent_t> uploaded(tiles.size());
    for (size_t t = 0; t < tiles.size(); ++t) {   // load_dense: the reference's dpu_broadcast_to (spmm_mul_csr.c:352-367)
        const Tile &tl = tiles[t];
        if (rowsB)
            CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(g->d_B) + tl.dev_off, tl.width * s, tl.host, tl.host_ld * s,
                                       tl.width * s, rowsB, cudaMemcpyHostToDevice, sin));
        uploaded[t] = new_event();
        CUDA_TRY(cudaEventRecord(uploaded[t], sin));
    }
    if (tiles.empty() || rowsC == 0) {
        CUDA_TRY(cudaEventRecord(g->ev[1], st));
        CUDA_TRY(cudaEventRecord(g->ev[2], st));
    }
    for (size_t t = 0; t < tiles.size() && rowsC; ++t) {
        const Tile &tl = tiles[t];
        CUDA_TRY(cudaStreamWaitEvent(st, uploaded[t], 0));
        if (t == 0) CUDA_TRY(cudaEventRecord(g->ev[1], st));          // first tile on the device: kernels start
        const bool last = t + 1 == tiles.size();
        // the last TWO tiles are cut into row chunks: the download stream must be idle when the last tile's chunks
        // arrive, so the tile before it has to start downloading while it still computes
        const int chunks_here = (chunk_last && t + 2 >= tiles.size()) ? n_chunks : 1;
        for (int k = 0; k < chunks_here; ++k) {
            long long brow = 0;
            for (size_t i = 0; i < g->parts.size(); ++i) {            // sparse part 0 overwrites, parts >= 1 add
                SparsePart &p = g->parts[i];
                const char *Bt = static_cast<const char *>(g->d_B) + tl.dev_off + (size_t)brow * tl.width * s;
                char *Ct = static_cast<char *>(g->d_C) + tl.col0 * s;
                int rc = run_tile(g, p, Bt, (long long)tl.width, Ct, (long long)H, (long long)tl.width, i > 0, st, nullptr,
                                  chunks_here > 1 ? &p.chunks[k] : nullptr);
                if (rc) return rc;
                brow += p.ncols;
            }
            size_t r0 = 0, nr = rowsC;
            if (chunks_here > 1) {
                r0 = (size_t)g->parts[0].chunks[k].row_begin;
                nr = (size_t)(g->parts[0].chunks[k].row_end - g->parts[0].chunks[k].row_begin);
            }
            if (last && k + 1 == chunks_here) CUDA_TRY(cudaEventRecord(g->ev[2], st));   // kernels done
            cudaEvent_t done = new_event();
            CUDA_TRY(cudaEventRecord(done, st));
            CUDA_TRY(cudaStreamWaitEvent(sout, done, 0));
            if (nr)   // retrieve_result (spmm_mul_csr.c:385-410); no merge step follows
                CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(C) + (r0 * (size_t)ldc + tl.col0) * s, (size_t)ldc * s,
                                           static_cast<char *>(g->d_C) + (r0 * H + tl.col0) * s, H * s, tl.width * s, nr,
                                           cudaMemcpyDeviceToHost, sout));
        }
    }
    return PYGIM_OK;
}

// join the download stream into the compute stream, wait, and read the phase timers of `g` (ev[3] = last download done)
static int host_join(Group *g, cudaStream_t st, cudaStream_t sout) {
    {
        cudaEvent_t e = g->chunk_done.back();          // host_enqueue keeps spare events at the end
        CUDA_TRY(cudaEventRecord(e, sout));
        CUDA_TRY(cudaStreamWaitEvent(st, e, 0));
        CUDA_TRY(cudaEventRecord(g->ev[3], st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    float ms = 0;
    g->timers_ms[0] = 0;   // load_sparse: done once in to_device_group
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev[0], g->ev[1])); g->timers_ms[1] = ms;   // exposed part of the upload
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev[1], g->ev[2])); g->timers_ms[2] = ms;   // kernels (uploads/downloads overlap)
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev[2], g->ev[3])); g->timers_ms[3] = ms;   // exposed tail of the download
    g->timers_ms[4] = 0;   // alignment: none
    return PYGIM_OK;
}

static int host_streams(Group *g) {
    if (!g->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
    if (!g->copy_in_stream) CUDA_TRY(cudaStreamCreateWithFlags(&g->copy_in_stream, cudaStreamNonBlocking));
    return PYGIM_OK;
}

PYGIM_API int pygim_spmm_run_group_host(pygim_handle_t handle, int n_ds, const void *const *B_parts, const int64_t *ldb,
                              void *C, int64_t ldc) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!B_parts || !ldb || !C) return fail(PYGIM_ERR_INVALID, "null buffer");
    CUDA_TRY(cudaSetDevice(g->device));
    int rc = host_streams(g);
    if (rc) return rc;
    cudaStream_t st = cudaStreamPerThread;
    rc = host_enqueue(g, n_ds, B_parts, ldb, C, ldc, st, g->copy_in_stream, g->copy_stream, true, true);
    if (rc) {
        (void)cudaStreamSynchronize(g->copy_in_stream);   // never return with copies from the caller's buffers in flight
        (void)cudaStreamSynchronize(g->copy_stream);
        (void)cudaStreamSynchronize(st);
        return rc;
    }
    return host_join(g, st, g->copy_stream);
}

PYGIM_API int pygim_spmm_run_many_host(int n_calls, const pygim_handle_t *handles, const void *const *B, const int64_t *ldb,
                                       void *const *C, const int64_t *ldc) {
    if (n_calls <= 0) return PYGIM_OK;
    if (!handles || !B || !ldb || !C || !ldc) return fail(PYGIM_ERR_INVALID, "null argument");
    std::vector<Group *> gs((size_t)n_calls);
    for (int k = 0; k < n_calls; ++k) {
        gs[(size_t)k] = as_group(handles[k]);
        if (!gs[(size_t)k]) return PYGIM_ERR_INVALID;
        if (!B[k] || !C[k]) return fail(PYGIM_ERR_INVALID, "null buffer (call %d)", k);
        if (gs[(size_t)k]->device != gs[0]->device) return fail(PYGIM_ERR_INVALID, "all plans of a batch must live on one device");
        for (int j = 0; j < k; ++j)
            if (gs[(size_t)j] == gs[(size_t)k]) return fail(PYGIM_ERR_INVALID, "a plan may appear once per batch (its staging buffers are per plan)");
    }
    CUDA_TRY(cudaSetDevice(gs[0]->device));
    int rc = host_streams(gs[0]);
    if (rc) return rc;
    // ONE upload stream and ONE download stream for the whole batch: PCIe transfers run in the order given
    cudaStream_t st = cudaStreamPerThread, sin = gs[0]->copy_in_stream, sout = gs[0]->copy_stream;
    for (int k = 0; k < n_calls && !rc; ++k) {
        Group *g = gs[(size_t)k];
        std::vector<const void *> parts;
        std::vector<long long> lds;
        column_tiles(g, B[k], ldb[k], parts, lds);
        std::vector<int64_t> l64(lds.begin(), lds.end());
        rc = host_enqueue(g, (int)parts.size(), parts.data(), l64.data(), C[k], ldc[k], st, sin, sout, k == 0, k + 1 == n_calls);
    }
    if (rc) {
        (void)cudaStreamSynchronize(sin);
        (void)cudaStreamSynchronize(sout);
        (void)cudaStreamSynchronize(st);
        return rc;
    }
    rc = host_join(gs[(size_t)n_calls - 1], st, sout);
    if (rc) return rc;
    // phase timers of the earlier calls: their events have all completed
    for (int k = 0; k + 1 < n_calls; ++k) {
        Group *g = gs[(size_t)k];
        float ms = 0;
        g->timers_ms[0] = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, g->ev[0], g->ev[1])); g->timers_ms[1] = ms;
        CUDA_TRY(cudaEventElapsedTime(&ms, g->ev[1], g->ev[2])); g->timers_ms[2] = ms;
        g->timers_ms[3] = 0;
        g->timers_ms[4] = 0;
    }
    return PYGIM_OK;
}

PYGIM_API int pygim_last_timers(pygim_handle_t handle, double *out5_ms) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!out5_ms) return fail(PYGIM_ERR_INVALID, "null output");
    for (int i = 0; i < 5; ++i) out5_ms[i] = g->timers_ms[i];
    return PYGIM_OK;
}

PYGIM_API int pygim_last_launches(pygim_handle_t handle, int64_t *out) {
    Group *g = as_group(handle);
    if (!g) return PYGIM_ERR_INVALID;
    if (!out) return fail(PYGIM_ERR_INVALID, "null output");
    *out = g->last_launches;
    return PYGIM_OK;
}

PYGIM_API int pygim_partition_rows_by_nnz(const int32_t *rowptr, int64_t nrows, int nparts, int64_t *split_out) {
    if (!rowptr || !split_out || nparts <= 0 || nrows < 0) return fail(PYGIM_ERR_INVALID, "bad partition arguments");
    auto at = [&](long long r) { return (long long)(unsigned)rowptr[r]; };
    const long long nnz = at(nrows);
    // Contiguous partition with the smallest possible heaviest part: binary search on the bound T, feasibility by
    // a greedy sweep that jumps with upper_bound over the prefix sums (rowptr itself).
    auto sweep = [&](long long T, int64_t *out) -> int {      // number of parts needed with every part <= T
        long long r = 0;
        int parts = 0;
        while (r < nrows) {
            const long long limit = at(r) + T;
            const int32_t *it = std::upper_bound(rowptr + r, rowptr + nrows + 1, limit,
                                                 [](long long t, int32_t a) { return t < (long long)(unsigned)a; });
            long long next = (it - rowptr) - 1;                // last boundary with prefix <= limit
            if (next <= r) next = r + 1;                       // a single row heavier than T (only when T < max row)
            if (out && parts + 1 <= nparts) out[parts + 1] = next;
            ++parts;
            r = next;
        }
        return parts;
    };
    long long lo = 0, hi = std::max<long long>(nnz, 1);
    for (long long r = 0; r < nrows; ++r) lo = std::max(lo, at(r + 1) - at(r));
    lo = std::max(lo, (nnz + nparts - 1) / nparts);
    while (lo < hi) {
        const long long mid = lo + (hi - lo) / 2;
        if (sweep(mid, nullptr) <= nparts) hi = mid; else lo = mid + 1;
    }
    split_out[0] = 0;
    for (int p = 1; p <= nparts; ++p) split_out[p] = nrows;    // parts the sweep does not need stay empty
    sweep(lo, split_out);
    split_out[nparts] = nrows;
    return PYGIM_OK;
}

PYGIM_API int pygim_partition_rows_even(int64_t nrows, int nparts, int64_t *split_out) {
    if (!split_out || nparts <= 0 || nrows < 0) return fail(PYGIM_ERR_INVALID, "bad partition arguments");
    const long long chunk = nrows / nparts, rest = nrows % nparts;
    long long cur = 0;
    split_out[0] = 0;
    for (int i = 0; i < nparts; ++i) {
        cur += chunk + (i < rest ? 1 : 0);
        split_out[i + 1] = cur;
    }
    return PYGIM_OK;
}

}  // extern "C"
