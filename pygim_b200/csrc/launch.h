// Type-erased launch descriptors shared by backend_pim.cu (C ABI, plans) and kernels_inst.cu
// (one translation unit per dtype, compiled with -DPYGIM_T=... -DPYGIM_SFX=...).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pygim {

struct Seg;

struct CsrLaunch {
    const int *rowptr;
    const int *colind;
    const void *val;
    const void *B;
    void *C;
    void *partial;            // scratch [n_seg x ldp]
    const Seg *segs;
    const int *long_rows;
    const int *long_seg_ptr;
    int *seg_count;           // [ceil(ncols/32) x n_long] zeroed arrival counters of the long rows
    int n_seg, n_long, nrows, seg_len;
    int rows_per_ticket;      // consecutive rows one work ticket covers (short-row graphs)
    int short_rows;           // mean degree is small: 1 = high-occupancy instantiation, 2 = + streamed row tickets
    long long ncols;          // dense columns of this tile
    long long ldb, ldc, ldp;  // row strides in elements
    int accumulate;
    int unit_values;          // every stored value is one: skip the value stream (bit-identical result)
    // fused all-gather: n_peers > 0 => rows go to peers[p] + (same offset as C) for every p (C is then unused);
    // mc != NULL => an NVSwitch multicast mapping of the same buffers (one multimem.st instead of n_peers stores)
    void *peers[8];
    void *mc;
    int n_peers;
    int sm_count;
    unsigned long long *ticket;        // two device counters of the plan (work tickets, warps out), zero at rest
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
    int n_warp_slots;         // resident warps of the device (for the automatic chunk size)
    int sm_count;
    unsigned long long *ticket;
    cudaStream_t stream;
};

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
