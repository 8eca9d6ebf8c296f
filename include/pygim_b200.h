/* pygim_b200 - C ABI of libbackend_pim.so (sm_100a).
 *
 * This is the drop-in boundary for PyGim's aggregation path (sparse adjacency x dense
 * features).  In the reference the boundary is a TORCH_LIBRARY(pim_ops, ...) block that is
 * dlopen()ed through torch.ops.load_library(args.lib_path) (spmm_test.py:111,
 * inference.py:134); one such library exists per (variant, dtype, format, balance) build.
 * Here there is ONE library, dtype/format are run-time arguments, and the entry points take
 * plain pointers and sizes so that any host language can bind them (ctypes in
 * pygim_b200/_lib.py; INTEGRATION.md shows the stub a PyGim maintainer would add).
 *
 * Every function returns 0 on success and a non-zero pygim_status_t otherwise;
 * pygim_last_error() returns the message for the calling thread.  The reference's failure
 * mode is assert()/DPU_ASSERT -> abort (spmm_default/spmm_mul_csr.c:136-137,251); here
 * errors are reported to the caller instead.
 *
 * All paths below are relative to /root/reference/backend_pim/.
 */
#ifndef PYGIM_B200_H
#define PYGIM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PYGIM_API __attribute__((visibility("default")))
#else
#define PYGIM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* val_dt, selected at compile time in the reference: spmm_default/support/common.h:39-60 */
typedef enum {
    PYGIM_INT8 = 0,
    PYGIM_INT16 = 1,
    PYGIM_INT32 = 2,
    PYGIM_INT64 = 3,
    PYGIM_FLT32 = 4,
    PYGIM_DBL64 = 5
} pygim_dtype_t;

/* KERNEL=MUL_CSR|MUL_COO in the reference build: spmm_default/CMakeLists.txt:28-30,53-60 */
typedef enum { PYGIM_CSR = 0, PYGIM_COO = 1 } pygim_format_t;

/* where the caller's buffers live */
typedef enum { PYGIM_MEM_HOST = 0, PYGIM_MEM_DEVICE = 1 } pygim_mem_t;

typedef enum {
    PYGIM_OK = 0,
    PYGIM_ERR_INVALID = 1,   /* bad argument (the reference assert()s: pytorch_api.cpp:212,252,264-266) */
    PYGIM_ERR_CUDA = 2,      /* a CUDA runtime call failed (the reference DPU_ASSERT -> abort) */
    PYGIM_ERR_NOT_INIT = 3,  /* dpu_init_* was not called (the reference dereferences a null dpu set) */
    PYGIM_ERR_NO_DEVICE = 4  /* no sm_100 device visible; there is NO CPU fallback */
} pygim_status_t;

typedef uint64_t pygim_handle_t;

/* ---------------------------------------------------------------- diagnostics */
PYGIM_API const char *pygim_last_error(void);
PYGIM_API int pygim_abi_version(void);

/* ---------------------------------------------------------------- device bring-up
 * Replaces dpu_init_ranks / dpu_init_dpus / dpu_release
 *   spmm_default/pytorch_api.cpp:154-164, spmm_grande/pytorch_api.cpp:157-181 (returns dpus_per_rank),
 *   spmm_multigroup/pytorch_api.cpp:168-178 (groups_per_rank).
 * A "rank" maps to one (sparse part, dense part) tile of the 2-D partitioning; on the GPU it
 * is a scheduling unit only.  `device` < 0 keeps the current CUDA device.
 * units_per_rank_out (may be NULL) receives nr_ranks entries - the analogue of grande's
 * dpus_per_rank list that grande.py:63-72 uses to deal the feature columns; each entry is the
 * number of column slices a rank is dealt (= SM count / nr_ranks, at least 1). */
PYGIM_API int pygim_dpu_init_ranks(int64_t nr_ranks, int64_t groups_per_rank, int device, int32_t *units_per_rank_out);
PYGIM_API int pygim_dpu_init_dpus(int64_t nr_dpus, int device);
PYGIM_API int pygim_dpu_release(void);

/* sm count, L2 size, max persisting-L2 bytes, total HBM bytes of the active device */
PYGIM_API int pygim_device_info(int *device, int *sm_count, int64_t *l2_bytes, int64_t *persisting_l2_max_bytes,
                      int64_t *hbm_bytes, int *cc_major, int *cc_minor);

/* ---------------------------------------------------------------- plan ("to_device_group")
 * Replaces spmm_csr_to_device_group / spmm_coo_to_device_group
 *   spmm_default/pytorch_api.cpp:204-243, :286-329 -> ops.hpp:65-92, :121-149
 *   -> prepare_pim_csr (spmm_mul_csr.c:118-259) / prepare_pim_coo (spmm_mul_coo.c:83-249)
 *   -> copy_sparse_csr (:261-330) / copy_sparse_coo (spmm_mul_coo.c:251-318)
 * and spmv_coo_to_device_group (spmv_sparseP/pytorch_api.cpp:184-229).
 *
 * n_sp sparse parts (the col_split of spmm.py:127-136) share nrows[0] rows; part i has ncols[i]
 * columns and nnz[i] nonzeros.  rowidx[i] is the CSR row pointer (nrows[i]+1 int32) or the COO row
 * index (nnz[i] int32, row-major sorted and coalesced as spmm.py:40-42 produces); colind[i] is
 * int32[nnz[i]]; values[i] has nnz[i] elements of `dtype`.  dense_cols[n_ds] are the widths of the
 * dense column parts (dense_split, spmm.py:9-13); they must sum to h_size.
 *
 * mem == PYGIM_MEM_HOST: arrays are copied to the device once, here (the reference uploads the
 * sparse parts once in copy_sparse_*).  mem == PYGIM_MEM_DEVICE: device arrays are BORROWED - the
 * caller keeps them alive until pygim_spmm_free_group (the reference borrows data_ptr()s the same
 * way, pytorch_api.cpp:230-232).
 *
 * The plan also holds the GPU analogue of the reference's two-level balancing
 * (support/partition.c:14-317): long rows are cut into nnz-bounded segments so that no warp
 * owns more than seg_len nonzeros. */
PYGIM_API int pygim_spmm_to_device_group(int format, int dtype, int n_sp, const int32_t *const *rowidx,
                               const int32_t *const *colind, const void *const *values, const int64_t *nrows,
                               const int64_t *ncols, const int64_t *nnz, int n_ds, const int64_t *dense_cols,
                               int64_t h_size, int mem, pygim_handle_t *out_handle);

/* spmm_free_group (spmm_default/pytorch_api.cpp:198-201; never called by the reference's Python) */
PYGIM_API int pygim_spmm_free_group(pygim_handle_t handle);

/* plan options (the knobs the retargeted autotuner turns, utils/autotuner.py); value < 0 = automatic:
 *   seg_len          rows with more nonzeros are cut into segments of at most this many
 *   item_nnz         short rows are grouped into work items of about this many nonzeros (default 256) ...
 *   rows_per_ticket  ... and at most this many rows (1..31)
 *   super_nnz        nonzeros per superticket - the unit an SM's warps drain together (default nnz / (16 SMs))
 *   cta_threads      threads per block (256 default; 1024 = ONE block per SM, every warp of the SM on the same
 *                    superticket: use with a locality-preserving row order so gathered rows are L1 hits)
 *   max_g            lanes per dense row are capped at this power of two; wider rows run as column chunks of one
 *                    launch (8 = 128-byte chunks, the L1 line)
 *   short_rows       CSR kernel family: 0 deep (128 registers, 16 gathers per lane in flight), 1 high occupancy
 *                    (40 registers), 2 streamed row items, 3 light (64 registers, 8 gathers), 4 two launches - rows of
 *                    at most 8 nonzeros by a matrix-wide grid of lane groups, the rest as <= seg_len pieces, longest
 *                    first; automatic: 4 when the mean degree is below 12, 3 below 96, else 0.  The family shapes
 *                    the plan: setting it re-plans
 *   unit_values      0 forces the general kernels even when every stored value is one
 *   coo_native       1 runs a sorted COO plan through the COO (segmented reduction) kernel instead of the CSR
 *                    kernels over the derived row pointer
 *   chunk_nnz        COO kernel: nonzeros per work item
 *   host_chunks      row chunks of the host entry point (0 = no download/compute overlap)
 *   host_tile_bytes  bytes per dense row of one upload / compute / download column tile of the host entry points
 *                    (multiple of 128; default 256: narrower 2-D PCIe copies lose a third of the rate when both
 *                    directions are busy)
 *   l2_persist       1/0 forces / forbids the access-policy window (persisting L2 lines) over the dense tile of a
 *                    launch; automatic = on when the tile fits the carve-out and every feature row is gathered >= 200
 *                    times per launch (it halves a long launch's DRAM traffic at equal time, but the pinned lines
 *                    slow down short launches that alternate between operands: 1/8 Reddit-shape shard sweep 1166 vs
 *                    913 us) */
PYGIM_API int pygim_plan_set_option(pygim_handle_t handle, const char *key, int64_t value);

/* Graph statistics the retargeted autotuner consumes (the role of pim_ops.prepare_tune_csr,
 * utils/autotuner.py:295-302,351): out[0..7] = nrows, ncols, nnz, max row nnz, number of long
 * rows (cut into segments), number of segments, seg_len, empty rows - for sparse part `part`. */
PYGIM_API int pygim_plan_stats(pygim_handle_t handle, int part, int64_t *out8);
/* out[0..5] = work items, supertickets, 1 if a COO plan runs through the CSR kernels (row-major sorted stream),
 * 1 if the COO stream is sorted, 1 if a row map is set, 1 if every stored value is one - for sparse part `part`. */
PYGIM_API int pygim_plan_layout(pygim_handle_t handle, int part, int64_t *out6);

/* Row reordering (the role ClusterData plays at spmm_test.py:57-65): the plan's sparse parts were
 * built from a ROW-PERMUTED adjacency (rows that share neighbours next to each other, so the SM that processes them
 * re-uses the gathered feature rows out of its L1); plan row r is row row_map[r] of the result, and the kernels
 * scatter on store - callers see the original row order.  n must equal the plan's row count; NULL clears the map. */
PYGIM_API int pygim_plan_set_row_map(pygim_handle_t handle, const int32_t *row_map, int64_t n, int mem);

/* ---------------------------------------------------------------- run ("run_group")
 * Replaces spmm_csr_run_group / spmm_coo_run_group / spmv_coo_run_group
 *   spmm_default/pytorch_api.cpp:248-280, :332-367 -> spmm_pim_csr (spmm_mul_csr.c:335-561),
 *   spmm_pim_coo (spmm_mul_coo.c:323-592); spmv_sparseP/pytorch_api.cpp:231-266.
 * Semantics (ops.hpp:42-62): C[total_rows x h_size] = sum_i concat_j A_i * B_j[rows_i, :], where
 * rows_i is the row block of B matching sparse part i's columns.  B_parts[j] is [sum_i ncols[i] x
 * dense_cols[j]] with row stride ldb[j] (elements); C has row stride ldc (elements).
 *
 * _host: B_parts and C are host buffers (pinned buffers transfer at full PCIe rate); the call
 *        uploads B, runs, downloads C and returns when C is complete - the reference's
 *        load_dense / kernel / retrieve_result / alignment phases.
 * _device: B_parts and C are device buffers; work is enqueued on `stream` (a cudaStream_t, NULL =
 *        legacy default stream) and the call returns without synchronising. */
PYGIM_API int pygim_spmm_run_group_host(pygim_handle_t handle, int n_ds, const void *const *B_parts, const int64_t *ldb,
                              void *C, int64_t ldc);
PYGIM_API int pygim_spmm_run_group_device(pygim_handle_t handle, int n_ds, const void *const *B_parts, const int64_t *ldb,
                                void *C, int64_t ldc, void *stream);

/* A BATCH of host-operand SpMMs - the hidden-size sweep of spmm_test.py:119-132 (one prepare + mul per dense size),
 * or the layers of inference.py:142-149 - as ONE software pipeline: call k runs handles[k] on the host matrix B[k]
 * ([sum ncols x h_size_k], row stride ldb[k] elements) into the host matrix C[k] (row stride ldc[k]).  All uploads
 * share one stream and run in the order given, tile by tile (256-byte column tiles); the kernels of a tile start as
 * soon as it has landed; every finished tile downloads while the next computes.  Per-call entry points expose the
 * first upload and the last download of EVERY call; the batch exposes one tile's upload and a quarter of one tile's
 * download in total.  Order the calls by ascending operand size so the kernels start early.  A plan may appear once
 * per batch.  Returns when every C[k] is complete; pygim_last_timers reports per-call phases as before. */
PYGIM_API int pygim_spmm_run_many_host(int n_calls, const pygim_handle_t *handles, const void *const *B, const int64_t *ldb,
                                       void *const *C, const int64_t *ldc);

/* Same computation when the caller already has B as ONE [sum ncols x h_size] device matrix (no
 * dense_split copies): the dense column parts become column tiles of B/C. */
PYGIM_API int pygim_spmm_device(pygim_handle_t handle, const void *B, int64_t ldb, void *C, int64_t ldc, void *stream);

/* Row-sharded multi-GPU run with the all-gather FUSED into the kernel epilogue (no reference counterpart: the
 * reference gathers DPU row blocks with dpu_push_xfer(FROM_DPU) and merges them on the host,
 * spmm_default/spmm_mul_csr.c:385-410,479-554).  This rank's plan covers rows [row_offset, row_offset + nrows)
 * of the global result; every output row is stored to that row of EACH of the n_peers (<= 8) result matrices
 * C_peers[q] - NVLink peer mappings of the other ranks' buffers, the local buffer included - or, when
 * C_multicast is not NULL, once to an NVSwitch multicast mapping of them (multimem.st).  CSR, sp_parts == 1.
 * The caller provides the cross-rank barriers before (peers done reading) and after (rows visible) the call. */
PYGIM_API int pygim_spmm_device_peers(pygim_handle_t handle, const void *B, int64_t ldb, void *const *C_peers,
                                      int n_peers, void *C_multicast, int64_t ldc, int64_t row_offset, void *stream);

/* Hot/cold plan for graphs with community structure (csrc/spmm_csr_hc.cuh; built by pygim_b200/reorder.py): the
 * rows are cut into n_super supertickets at super_rows[0..n_super] (row ids, super_rows[0] = 0, last = nrows); for
 * superticket k the hot_k columns hot_cols[k * hot_k ...] (-1 = unused) are staged in shared memory while its rows
 * are processed.  The plan's colind must already be in hot/cold form: the first hot_cnt[r] nonzeros of row r hold
 * TILE SLOTS (0 .. hot_k-1), the rest column ids; rows longer than the plan's seg_len must be all cold.  Set seg_len
 * (and other item options) BEFORE this call.  sp_parts == 1. */
PYGIM_API int pygim_plan_set_hot_tiles(pygim_handle_t handle, int64_t n_super, const int32_t *super_rows, int hot_k,
                                       const int32_t *hot_cols, const int32_t *hot_cnt, int mem);

/* Everything a conv layer does around the aggregation, fused into the row store of the SpMM
 * (models/pyg_gcn_conv.py:130-137, pyg_gin_conv.py:80-101, models/quantize.py:40-42), plus the multi-GPU exchange.
 * All fields optional (zero = off):
 *   scale            device scalar: the result is FLOAT32, (float)sum * scale[0] - symmetric_dequantize
 *   residual         device float32 [rows x h_size], row stride ld_residual: result += residual_coeff * residual
 *                    (GIN's (1 + eps) * x_r); one multiply, one multiply, one add, never contracted
 *   C_peers / n_peers / C_multicast / row_offset   as pygim_spmm_device_peers (C is then ignored)
 *   row_peer_mask    device, one byte per plan row: bit p set = peer p needs this row (halo exchange); NULL = all
 *   flag_peers / my_rank / epoch   in-kernel arrival signal: when this rank's last row has been stored the kernel
 *                    writes `epoch` to flag_peers[p][my_rank] of every peer p (release, system scope); consumers
 *                    wait with pygim_wait_flags instead of a cross-rank barrier per call
 * ldc is in elements of the RESULT type (float32 when scale or residual is set).  Needs sp_parts == 1. */
typedef struct {
    const float *scale;
    const float *residual;
    int64_t ld_residual;
    float residual_coeff;
    void *const *C_peers;
    int n_peers;
    void *C_multicast;
    int64_t row_offset;
    const uint8_t *row_peer_mask;
    int32_t *const *flag_peers;
    int my_rank;
    int32_t epoch;
} pygim_epilogue_t;
PYGIM_API int pygim_spmm_device_ex(pygim_handle_t handle, const void *B, int64_t ldb, void *C, int64_t ldc,
                                   const pygim_epilogue_t *epi, void *stream);
/* enqueue a wait on `stream` until flags[0..n) have all reached `epoch` (n <= 32) */
PYGIM_API int pygim_wait_flags(const int32_t *flags, int n, int32_t epoch, void *stream);

/* symmetric_quantize (models/quantize.py:20-38) as two kernels: scale[0] = 2 * max|x| / 2^k (k = 5 / 10 / 20 for
 * INT8 / INT16 / INT32 and FLT32), xq = round_half_even(x / scale) stored as `dtype`.  x is float32
 * [rows x cols] with row stride ldx, xq has row stride ldq; scale is a device float.  Bit-identical to the torch
 * expression (IEEE division, rintf).  INT64 / DBL64 are rejected (the reference's else-branch quantises to float). */
PYGIM_API int pygim_quantize(const float *x, int64_t rows, int64_t cols, int64_t ldx, int dtype, void *xq, int64_t ldq,
                             float *scale, void *stream);

/* The five phase timers the reference prints as [DATA]load_sparse_time / load_dense_time /
 * kernel_time / retrieve_result_time / alignment_time (spmm_mul_csr.c:563-580), in ms, for the
 * last pygim_spmm_run_group_host call on this handle (alignment is always 0: the kernels write C
 * in place, there is no host merge). */
PYGIM_API int pygim_last_timers(pygim_handle_t handle, double *out5_ms);

/* number of kernel launches issued by the last run call on this handle */
PYGIM_API int pygim_last_launches(pygim_handle_t handle, int64_t *out);

/* ---------------------------------------------------------------- partitioners (host, no GPU needed)
 * GPU-level analogue of partition_by_nnz_csr (support/partition.c:51-99): cut [0, nrows) into
 * nparts contiguous row ranges of near-equal nnz.  Where the reference sweeps greedily against nnz/n
 * (and merges the leftovers into the last part), this returns the contiguous partition whose HEAVIEST part is as
 * light as possible (binary search on the bound + greedy sweep), so it is never worse balanced than the
 * reference's.  split_out has nparts+1 entries; parts that are not needed are empty. */
PYGIM_API int pygim_partition_rows_by_nnz(const int32_t *rowptr, int64_t nrows, int nparts, int64_t *split_out);
/* partition_by_row_csr (support/partition.c:14-44): nrows/nparts rows each, first nrows%nparts get +1 */
PYGIM_API int pygim_partition_rows_even(int64_t nrows, int nparts, int64_t *split_out);

#ifdef __cplusplus
}
#endif
#endif /* PYGIM_B200_H */
