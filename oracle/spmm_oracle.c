/* TEST INFRASTRUCTURE ONLY - the CPU oracle for pygim_b200.
 *
 * A plain-C restatement of the arithmetic PyGim defines for its aggregation
 * path C = A_sparse * B_dense.  Nothing in the product path (pygim_b200/) may
 * import, link or execute this file; it exists so that tests/, smoke() and
 * bench.py's cpu_baseline leg have something to check and time against.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every
 * function here against the reference's own host oracles compiled in place
 * from /root/reference (oracle/build_ref.sh -> oracle/_ref/), and against the
 * golden vectors under tests/golden/ that were generated from those same
 * reference objects (tests/golden/make_golden.py).
 *
 * Reference anchors (paths relative to /root/reference/backend_pim):
 *   COO definition        spmm_default/spmm_mul_coo.c:40-51   (y[r*H+k] += x[c*H+k]*val, nnz order, in val_dt)
 *   CSR definition        spmm_grande/spmm_mul_csr.c:119-136  (uses values; x has a padded row stride)
 *   CSR, values ignored   spmm_default/spmm_mul_csr.c:100-113 (kept as *_csr_ones: "Assuming that values are 1s")
 *   SpMV flavour          spmv_sparseP/spmv_mul_coo.c:92-103
 *   group composition     spmm_default/ops.hpp:42-62,97-118   (sum over sparse parts, concat over dense parts)
 *   tile placement        spmm_default/spmm_mul_csr.c:41-86   (add_2D / memcpy_2D / memadd_2D)
 *   accumulator width     spmm_default/dpu_kernels/spmm_mul_csr_dpu.c:72-75,110-114 (acc is val_dt => ints wrap)
 *   dtype table           spmm_default/support/common.h:39-60
 *
 * Integer types: the reference accumulates in val_dt, i.e. modulo 2^bits.  We
 * do the arithmetic in the unsigned type of the same width, which is the
 * defined-behaviour spelling of that wraparound.  Floating types: one multiply
 * and one add per nonzero, in nnz order, no contraction (built with
 * -ffp-contract=off) so the result is a deterministic function of the inputs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* T = storage type, U = arithmetic type (unsigned twin for integers) */
#define DEFINE_ORACLE(SUFFIX, T, U)                                                                   \
    /* spmm_default/spmm_mul_coo.c:40-51 */                                                           \
    ORACLE_API void oracle_spmm_coo_##SUFFIX(T *y, int64_t nnz, const int32_t *rowind,                \
                                             const int32_t *colind, const T *val, const T *x,         \
                                             int64_t ncols) {                                         \
        for (int64_t n = 0; n < nnz; n++) {                                                           \
            int64_t r = (uint32_t)rowind[n], c = (uint32_t)colind[n];                                 \
            U v = (U)val[n];                                                                          \
            for (int64_t k = 0; k < ncols; k++)                                                       \
                y[r * ncols + k] = (T)((U)y[r * ncols + k] + (U)((U)x[c * ncols + k] * v));           \
        }                                                                                             \
    }                                                                                                 \
    /* spmm_grande/spmm_mul_csr.c:119-136: column-outer loop, x row stride = ncols_pad */             \
    ORACLE_API void oracle_spmm_csr_##SUFFIX(T *y, int64_t nrows, const int32_t *rowptr,              \
                                             const int32_t *colind, const T *values, const T *x,      \
                                             int64_t ncols, int64_t ncols_pad) {                      \
        for (int64_t r = 0; r < nrows; r++)                                                           \
            for (int64_t k = 0; k < ncols; k++)                                                       \
                for (int64_t i = (uint32_t)rowptr[r]; i < (int64_t)(uint32_t)rowptr[r + 1]; i++) {    \
                    int64_t c = (uint32_t)colind[i];                                                  \
                    y[r * ncols + k] =                                                                \
                        (T)((U)y[r * ncols + k] + (U)((U)values[i] * (U)x[c * ncols_pad + k]));       \
                }                                                                                     \
    }                                                                                                 \
    /* spmm_default/spmm_mul_csr.c:100-113: same walk, values NOT used */                             \
    ORACLE_API void oracle_spmm_csr_ones_##SUFFIX(T *y, int64_t nrows, const int32_t *rowptr,         \
                                                  const int32_t *colind, const T *x, int64_t ncols) { \
        for (int64_t r = 0; r < nrows; r++)                                                           \
            for (int64_t k = 0; k < ncols; k++)                                                       \
                for (int64_t i = (uint32_t)rowptr[r]; i < (int64_t)(uint32_t)rowptr[r + 1]; i++) {    \
                    int64_t c = (uint32_t)colind[i];                                                  \
                    y[r * ncols + k] = (T)((U)y[r * ncols + k] + (U)x[c * ncols + k]);                \
                }                                                                                     \
    }                                                                                                 \
    /* spmm_default/spmm_mul_csr.c (add_2D): A[off_x+i][off_y+j] += B[i][j] */                        \
    ORACLE_API void oracle_add_2d_##SUFFIX(T *A, const T *B, int64_t A_ncols, int64_t B_ncols,        \
                                           int64_t off_x, int64_t off_y, int64_t len_x,               \
                                           int64_t len_y) {                                           \
        for (int64_t i = 0; i < len_x; i++)                                                           \
            for (int64_t j = 0; j < len_y; j++)                                                       \
                A[(off_x + i) * A_ncols + off_y + j] =                                                \
                    (T)((U)A[(off_x + i) * A_ncols + off_y + j] + (U)B[i * B_ncols + j]);             \
    }                                                                                                 \
    /* spmm_default/ops.hpp:42-62 (CSR) and :97-118 (COO): y = sum_i ( concat_j A_i * B_j[rows_i] ). \
     * fmt 0 = CSR (values used), 1 = COO.  B_parts[j] is [total_Brows x h_j] contiguous; sparse    \
     * part i multiplies rows [current_Brow, current_Brow + ncols_i) of every B part.  y must be     \
     * zero-initialised by the caller (torch::zeros, pytorch_api.cpp:270-271). */                     \
    ORACLE_API void oracle_spmm_group_##SUFFIX(T *y, int fmt, int n_sp, const int64_t *nrows,         \
                                               const int64_t *ncols_sp, const int64_t *nnz,           \
                                               const int32_t *const *rowidx,                          \
                                               const int32_t *const *colind, const T *const *values,  \
                                               int n_ds, const T *const *B_parts,                     \
                                               const int64_t *h_sizes, int64_t total_cols) {          \
        int64_t current_Brow = 0;                                                                     \
        for (int i = 0; i < n_sp; i++) {                                                              \
            int64_t max_h = 0;                                                                        \
            for (int j = 0; j < n_ds; j++)                                                            \
                if (h_sizes[j] > max_h) max_h = h_sizes[j];                                           \
            T *y_temp = (T *)malloc(sizeof(T) * (size_t)(max_h > 0 ? max_h : 1) *                     \
                                    (size_t)(nrows[i] > 0 ? nrows[i] : 1));                           \
            int64_t current_Acol = 0;                                                                 \
            for (int j = 0; j < n_ds; j++) {                                                          \
                int64_t h = h_sizes[j];                                                               \
                memset(y_temp, 0, sizeof(T) * (size_t)h * (size_t)nrows[i]);                          \
                const T *x = B_parts[j] + current_Brow * h;                                           \
                if (fmt == 0)                                                                         \
                    oracle_spmm_csr_##SUFFIX(y_temp, nrows[i], rowidx[i], colind[i], values[i], x, h, \
                                             h);                                                      \
                else                                                                                  \
                    oracle_spmm_coo_##SUFFIX(y_temp, nnz[i], rowidx[i], colind[i], values[i], x, h);  \
                oracle_add_2d_##SUFFIX(y, y_temp, total_cols, h, 0, current_Acol, nrows[i], h);       \
                current_Acol += h;                                                                    \
            }                                                                                         \
            current_Brow += ncols_sp[i];                                                              \
            free(y_temp);                                                                             \
        }                                                                                             \
    }                                                                                                 \
    /* The `--version=cpu` algorithm class (torch_sparse spmm_sum on CPU; call sites               \
     * spmm_test.py:25, models/pyg_gcn_conv.py:133): row-parallel CSR, per-row H-vector              \
     * accumulated in the element type, written once.  values == NULL means implicit ones           \
     * (SparseTensor value=None).  Strided so it can run on a row/column tile.  Used as the         \
     * timed CPU baseline and as the fast oracle for large cases. */                                  \
    ORACLE_API void oracle_spmm_csr_rowpar_##SUFFIX(T *y, int64_t ldy, int64_t nrows,                 \
                                                    const int32_t *rowptr, const int32_t *colind,     \
                                                    const T *values, const T *x, int64_t ldx,         \
                                                    int64_t ncols, int accumulate, int nthreads) {    \
        if (nthreads <= 0) nthreads = 1;                                                              \
        _Pragma("omp parallel num_threads(nthreads)") {                                               \
            U *acc = (U *)malloc(sizeof(U) * (size_t)(ncols > 0 ? ncols : 1));                        \
            _Pragma("omp for schedule(dynamic, 64)") for (int64_t r = 0; r < nrows; r++) {            \
                for (int64_t k = 0; k < ncols; k++) acc[k] = accumulate ? (U)y[r * ldy + k] : (U)0;   \
                for (int64_t i = (uint32_t)rowptr[r]; i < (int64_t)(uint32_t)rowptr[r + 1]; i++) {    \
                    const T *xr = x + (int64_t)(uint32_t)colind[i] * ldx;                             \
                    if (values) {                                                                     \
                        U v = (U)values[i];                                                           \
                        for (int64_t k = 0; k < ncols; k++) acc[k] = (U)(acc[k] + (U)((U)xr[k] * v)); \
                    } else {                                                                          \
                        for (int64_t k = 0; k < ncols; k++) acc[k] = (U)(acc[k] + (U)xr[k]);          \
                    }                                                                                 \
                }                                                                                     \
                for (int64_t k = 0; k < ncols; k++) y[r * ldy + k] = (T)acc[k];                       \
            }                                                                                         \
            free(acc);                                                                                \
        }                                                                                             \
    }

DEFINE_ORACLE(i8, int8_t, uint8_t)
DEFINE_ORACLE(i16, int16_t, uint16_t)
DEFINE_ORACLE(i32, int32_t, uint32_t)
DEFINE_ORACLE(i64, int64_t, uint64_t)
DEFINE_ORACLE(f32, float, float)
DEFINE_ORACLE(f64, double, double)

/* f64-accumulated FLT32 SpMM plus the per-element magnitude sum_e |a_e*x_e| that the
 * float tolerance in tests/ is stated against:
 *     |gpu - exact| <= 1e-5 * mag + 1e-6.                                              */
ORACLE_API void oracle_spmm_csr_f32_exact(double *y, double *mag, int64_t nrows, const int32_t *rowptr,
                                          const int32_t *colind, const float *values, const float *x,
                                          int64_t ldx, int64_t ncols, int nthreads) {
    if (nthreads <= 0) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t r = 0; r < nrows; r++) {
        for (int64_t k = 0; k < ncols; k++) {
            y[r * ncols + k] = 0.0;
            mag[r * ncols + k] = 0.0;
        }
        for (int64_t i = (uint32_t)rowptr[r]; i < (int64_t)(uint32_t)rowptr[r + 1]; i++) {
            const float *xr = x + (int64_t)(uint32_t)colind[i] * ldx;
            double v = values ? (double)values[i] : 1.0;
            for (int64_t k = 0; k < ncols; k++) {
                double p = v * (double)xr[k];
                y[r * ncols + k] += p;
                mag[r * ncols + k] += p < 0 ? -p : p;
            }
        }
    }
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
