// Type-erased launch descriptors shared by backend_pim.cu (C ABI, plans) and kernels_inst.cu
// (one translation unit per dtype, compiled with -DPYGIM_T=... -DPYGIM_SFX=...).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pygim {

struct Seg;

// optional work of the row store (spmm_csr.cuh: struct Epilogue)
struct EpilogueLaunch {
    const int *row_map = nullptr;
    const float *scale = nullptr;
    const float *residual = nullptr;
    long long ld_res = 0;
    float coeff = 0.f;
    void *peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void *mc = nullptr;
    const unsigned char *peer_mask = nullptr;
    int n_peers = 0;
    int *flags[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int my_rank = 0;
    int epoch = 0;
};

struct CsrLaunch {
    const int *rowptr;
    const int *colind;
    const void *val;
    const void *B;
    void *C;                  // element type of the plan, or float32 when epi.scale / epi.residual is set
    void *partial;            // scratch [n_seg x ldp]
    const Seg *segs;
    const int4 *supers;       // [n_super] supertickets (x >= 0: rows [x, x+y), z rows per item, w items; x < 0: w segments from ~x)
    const int *long_rows;
    const int *long_seg_ptr;
    int *super_cnt;           // [n_super x col_chunks] zeroed draw counters
    int *seg_count;           // [col_chunks x n_long] zeroed arrival counters of the long rows
    unsigned int *warps_out;  // zeroed
    int n_super, n_items, n_seg, n_long, nrows, seg_len;
    long long nnz_total;
    int short_rows;           // mean degree is small: 1 = high-occupancy instantiation, 2 = + streamed row items
    int max_g;                // lanes per dense row are capped at this power of two (column chunks beyond it)
    int cta_threads;          // threads per block (256 .. 1024)
    long long ncols;          // dense columns of this tile
    long long ldb, ldc, ldp;  // row strides in elements (ldc in elements of the OUTPUT type)
    int accumulate;
    int unit_values;          // every stored value is one: skip the value stream (bit-identical result)
    // hot/cold plans (spmm_csr_hc.cuh): hot_k > 0 => colind holds tile slots for the first hot_cnt[r] nonzeros of row r
    const int *hot_cols = nullptr;   // [row supertickets x hot_k]
    const int *hot_cnt = nullptr;    // [nrows]
    int hot_k = 0;
    int n_seg_super = 0;
    EpilogueLaunch epi;
    int sm_count;
    cudaStream_t stream;
};

struct CooLaunch {
    const int *rowind;
    const int *colind;
    const void *val;
    const void *B;
    void *C;
    long long nnz, nrows, ncols;
    long long ldb, ldc;
    int chunk_nnz;            // target nonzeros per warp; <= 0 = automatic
    int unit_values;
    int accumulate;           // 0: the launcher zero-fills the C tile first
    int all_atomic;           // the stream is NOT row-major sorted: every flush is an atomic add
    int n_warp_slots;         // resident warps of the device (for the automatic chunk size)
    int sm_count;
    unsigned long long *ticket;
    cudaStream_t stream;
};

// ---- launch geometry shared by the launcher (kernels_inst.cu) and the scratch sizing (backend_pim.cu)
// 16-byte words when every row start is 16-byte aligned, single elements otherwise
inline bool csr_can_vectorize(size_t s, const void *B, const void *C, size_t out_elem, long long ncols, long long ldb,
                              long long ldc, long long ldp) {
    auto mis = [](const void *p) { return (int)(reinterpret_cast<uintptr_t>(p) & 15); };
    return mis(B) == 0 && mis(C) == 0 && (ncols * (long long)s) % 16 == 0 && (ldb * (long long)s) % 16 == 0 &&
           (ldc * (long long)out_elem) % 16 == 0 && (ldp * (long long)s) % 16 == 0;
}
inline int csr_lanes(long long nvec, int max_g) {      // lanes covering one dense row: power of two, <= max_g
    int g = 1;
    const int cap = max_g >= 1 && max_g <= 32 ? max_g : 32;
    while (g < nvec && g < cap) g <<= 1;
    return g;
}
inline int csr_col_chunks(size_t s, long long ncols, int max_g, bool vectorized) {
    const long long nvec = vectorized ? ncols * (long long)s / 16 : ncols;
    const int g = csr_lanes(nvec, max_g);
    return (int)((nvec + g - 1) / g);
}

#define PYGIM_DECLARE_LAUNCHERS(SFX)                                                                     \
    cudaError_t launch_csr_##SFX(const CsrLaunch &l, int64_t *launches);                                  \
    cudaError_t launch_coo_##SFX(const CooLaunch &l, int64_t *launches);                                  \
    cudaError_t check_all_ones_##SFX(const void *val, long long n, int *d_flag, cudaStream_t stream);

PYGIM_DECLARE_LAUNCHERS(i8)
PYGIM_DECLARE_LAUNCHERS(i16)
PYGIM_DECLARE_LAUNCHERS(i32)
PYGIM_DECLARE_LAUNCHERS(i64)
PYGIM_DECLARE_LAUNCHERS(f32)
PYGIM_DECLARE_LAUNCHERS(f64)

}  // namespace pygim
