/*
 * tabmat_b200 — C-ABI of the B200 (sm_100a) sandwich / matvec / transpose_matvec path.
 *
 * This header is the drop-in boundary.  Every entry point replaces one Python-callable
 * function of the reference's four Cython extension modules (the "plugin boundary",
 * SURVEY.md §8b); the reference function it stands in for is cited as file:line relative
 * to the reference tree (Quantco/tabmat @ 7af8b2c, src/tabmat/ext/...).
 *
 * Conventions (all entry points):
 *   - every data pointer is a DEVICE pointer on the current CUDA device, borrowed for
 *     the duration of the call; nothing is retained;
 *   - outputs are caller-allocated; "overwrites" means the function fully defines the
 *     output (no zero-initialisation needed), "accumulates" means `out +=`;
 *   - `rows` / `cols` are int32 index lists; NULL means "all rows/cols" (the reference
 *     materialises np.arange instead, util.py:6-24).  Lists must be sorted and unique
 *     (the reference silently assumes this);
 *   - `c_order` = 1 for row-major (C-contiguous), 0 for column-major (F-contiguous);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - return value 0 = ok, non-zero = error; tm_last_error() gives the message
 *     (the reference's native code has no error return; validation lives in Python);
 *   - scratch memory is taken from the stream-ordered device pool
 *     (cudaMallocAsync/cudaFreeAsync) — the B200 stand-in for alloc.h:35-54;
 *   - sparse index arrays are int32 (n, p, nnz < 2^31 per GPU shard).
 *   - `_f32` / `_f64` suffixes select the arithmetic type, mirroring the reference's
 *     Cython fused type `floating`.
 */
#ifndef TABMAT_B200_H
#define TABMAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tm_stream_t;

/* ---- library ---------------------------------------------------------------------- */
int tm_version(void);
const char* tm_last_error(void);
/* Number of kernel launches issued by this library since the last reset (bench.py's
 * `gpu_launches`). */
int64_t tm_launch_count(void);
void tm_reset_launch_count(void);
/* 1 if the tcgen05 (TMEM/TMA) dense path can run on the current device (cc 10.x). */
int tm_has_tcgen05(void);
/* Select the dense f32 sandwich implementation: 0 = auto (tcgen05 when eligible),
 * 1 = force the CUDA-core kernel, 2 = force tcgen05 (error when not eligible),
 * 3 = fp32-accurate: "3xTF32" on the tensor cores where the tcgen05 path applies (every operand
 * split into hi = tf32(a) and lo = tf32(a - hi), products hi*hi + hi*lo + lo*hi), the CUDA-core
 * kernel elsewhere.
 * NUMERICS: modes 0 and 2 round both operands of the dense f32 block (X and d*X) to TF32
 * (10-bit mantissa) before the MMA and accumulate in fp32: normwise error ~1e-4 of the largest
 * entry (inside BASELINE's 1e-3 fp32 tolerance, tests/test_gpu_baseline_configs.py prints it),
 * but small differences of large entries (centred moments of columns with a big mean) lose
 * ~3 digits against the reference's full-precision fp32 FMAs (dense_helpers-tmpl.cpp:97).
 * Mode 3 restores fp32-level accuracy at a third of the tensor throughput; mode 1 is exact
 * fp32.  Also settable with TABMAT_B200_DENSE_F32_MODE. */
void tm_set_dense_f32_mode(int mode);
/* Fused dense-operand cross pass (tm_dense_cross_sandwich_*, tm_split_sandwich_blocks_*):
 * 0 / 1 = run-aggregating kernel (default), 2 = the one-row-per-visit kernel. */
void tm_set_cross_runs_mode(int mode);
/* Number of SMs the one-wave gather kernel of the dense x sparse block leaves free, so that a
 * collective running on another stream (the allreduce of the blocks that are already finished,
 * tabmat_b200/distributed.py) has somewhere to run.  0 = none (default). */
void tm_set_sm_reserve(int sms);
/* The same for the persistent tcgen05 kernel: it runs on (SMs - sms) CTAs, so that a collective
 * started before it has SMs (and their shared memory) to run on.  0 = none (default). */
void tm_set_tc_sm_reserve(int sms);
/* SplitMatrix sandwich, f32: number of scatter warps appended to the tcgen05 kernel, which then
 * also computes the dense x many-level categorical blocks (run sums + vector REDs) - and, when
 * forced, the per-non-zero REDs of the dense x sparse block - from the TMA-staged tile.
 * -1 = auto (default: 4 warps when only categorical blocks are left to it, i.e. when the sparse
 * block goes through the gather form; none otherwise - with the sparse REDs the fused form
 * measured 28-52 ms against 18 ms for two passes), 0 = never, 4 or 8 = always;
 * other values are ignored. */
void tm_set_tc_scatter_warps(int warps);
/* fp32 -> tf32 rounding of the tcgen05 operands (round to nearest, ties away, like the
 * reference-free cvt.rna.tf32.f32): 0 = the cvt instruction (4 SASS instructions on sm_100a),
 * 1 = integer add + mask (same values for finite inputs and infinities), 2 = integer add only
 * (the tensor core ignores the low 13 bits of a tf32 operand; tested), -1 = default
 * (TABMAT_B200_TC_ROUND, else 2). */
void tm_set_tc_round_mode(int mode);
/* Length of an accumulation chain in the tcgen05 kernels, in K = 8 MMA steps.  The tensor core's
 * fp32 accumulator truncates at every step, so a chain of T steps onto a growing sum loses about
 * T * 2^-25 of it (measured -9.3e-4 on the diagonal of X^T D X at 4e7 rows on one GPU); a long
 * input is therefore cut into several launches, each draining its TMEM accumulators into the
 * result (fp32 REDs) after at most `steps` steps per CTA.
 * -1 = default (TABMAT_B200_TC_FLUSH_STEPS, else 4096 = a loss of 1.2e-4 .. 2.4e-4), 0 = one
 * launch whatever the length. */
void tm_set_tc_flush_steps(int steps);

/* ---- dense block (reference: ext/dense.pyx) --------------------------------------- */
/* dense_sandwich, dense.pyx:19-44 -> _dense{C,F}_sandwich, dense_helpers-tmpl.cpp:266-308.
 * out[a*n_cols+b] = sum_t X[rows[t], cols[a]] * d[rows[t]] * X[rows[t], cols[b]].
 * Overwrites out (n_cols x n_cols, row-major, exactly symmetric). */
int tm_dense_sandwich_f32(const float* X, int64_t n, int64_t p, int c_order, const float* d,
                          const int32_t* rows, int64_t n_rows, const int32_t* cols,
                          int64_t n_cols, float* out, tm_stream_t stream);
int tm_dense_sandwich_f64(const double* X, int64_t n, int64_t p, int c_order, const double* d,
                          const int32_t* rows, int64_t n_rows, const int32_t* cols,
                          int64_t n_cols, double* out, tm_stream_t stream);

/* dense_matvec, dense.pyx:76-101 -> _dense{C,F}_matvec, dense_helpers-tmpl.cpp:385-417.
 * res[r] = sum_c X[rows[r], cols[c]] * v[cols[c]]   (v indexed by absolute column).
 * accumulate=0: out[r] = res[r]; accumulate=1: out[r] += res[r]. */
int tm_dense_matvec_f32(const float* X, int64_t n, int64_t p, int c_order, const float* v,
                        const int32_t* rows, int64_t n_rows, const int32_t* cols,
                        int64_t n_cols, float* out, int accumulate, tm_stream_t stream);
int tm_dense_matvec_f64(const double* X, int64_t n, int64_t p, int c_order, const double* v,
                        const int32_t* rows, int64_t n_rows, const int32_t* cols,
                        int64_t n_cols, double* out, int accumulate, tm_stream_t stream);

/* dense_rmatvec, dense.pyx:48-73 -> _dense{C,F}_rmatvec, dense_helpers-tmpl.cpp:314-383.
 * out[c] = sum_t X[rows[t], cols[c]] * v[rows[t]].  Overwrites out (n_cols). */
int tm_dense_rmatvec_f32(const float* X, int64_t n, int64_t p, int c_order, const float* v,
                         const int32_t* rows, int64_t n_rows, const int32_t* cols,
                         int64_t n_cols, float* out, tm_stream_t stream);
int tm_dense_rmatvec_f64(const double* X, int64_t n, int64_t p, int c_order, const double* v,
                         const int32_t* rows, int64_t n_rows, const int32_t* cols,
                         int64_t n_cols, double* out, tm_stream_t stream);

/* transpose_square_dot_weights (dense), dense.pyx:103-122.
 * out[j] = sum_i w[i] * (X[i,j] - shift[j])^2.  Overwrites out (p). */
int tm_dense_sq_dot_weights_f32(const float* X, int64_t n, int64_t p, int c_order,
                                const float* w, const float* shift, float* out,
                                tm_stream_t stream);
int tm_dense_sq_dot_weights_f64(const double* X, int64_t n, int64_t p, int c_order,
                                const double* w, const double* shift, double* out,
                                tm_stream_t stream);

/* ---- sparse block (reference: ext/sparse.pyx) -------------------------------------- */
/* Device layout of a sparse block: CSR (data, indices, indptr[n+1]) with column indices
 * sorted inside each row, plus `csr_row` = the row id of every CSR non-zero (COO row
 * array, nnz entries); and CSC (data, indices, indptr[p+1]).  The reference keeps CSC
 * plus a lazily cached CSR (sparse_matrix.py:133-143). */

/* sparse_sandwich, sparse.pyx:17-77.  out = A[rows,cols]^T diag(d[rows]) A[rows,cols].
 * Overwrites out (n_cols x n_cols, exactly symmetric). */
int tm_sparse_sandwich_f32(const float* csr_data, const int32_t* csr_indices,
                           const int32_t* csr_indptr, const int32_t* csr_row, int64_t n,
                           int64_t p, int64_t nnz, const float* d, const int32_t* rows,
                           int64_t n_rows, const int32_t* cols, int64_t n_cols, float* out,
                           tm_stream_t stream);
int tm_sparse_sandwich_f64(const double* csr_data, const int32_t* csr_indices,
                           const int32_t* csr_indptr, const int32_t* csr_row, int64_t n,
                           int64_t p, int64_t nnz, const double* d, const int32_t* rows,
                           int64_t n_rows, const int32_t* cols, int64_t n_cols, double* out,
                           tm_stream_t stream);

/* csr_dense_sandwich, sparse.pyx:211-260 -> _csr_dense{C,F}_sandwich,
 * sparse_helpers-tmpl.cpp:23-143.
 * out[a*nB+b] = sum_t A[rows[t], A_cols[a]] * d[rows[t]] * B[rows[t], B_cols[b]].
 * Overwrites out (nA x nB). */
int tm_csr_dense_sandwich_f32(const float* csr_data, const int32_t* csr_indices,
                              const int32_t* csr_indptr, int64_t n, int64_t p_sparse,
                              const float* B, int64_t q, int b_c_order, const float* d,
                              const int32_t* rows, int64_t n_rows, const int32_t* A_cols,
                              int64_t nA, const int32_t* B_cols, int64_t nB, float* out,
                              tm_stream_t stream);
int tm_csr_dense_sandwich_f64(const double* csr_data, const int32_t* csr_indices,
                              const int32_t* csr_indptr, int64_t n, int64_t p_sparse,
                              const double* B, int64_t q, int b_c_order, const double* d,
                              const int32_t* rows, int64_t n_rows, const int32_t* A_cols,
                              int64_t nA, const int32_t* B_cols, int64_t nB, double* out,
                              tm_stream_t stream);

/* csr_matvec_unrestricted / csr_matvec, sparse.pyx:79-140.
 * res[t] = sum_{j in cols} X[rows[t], j] * v[j].  accumulate: out[t] (+)= res[t]. */
int tm_csr_matvec_f32(const float* csr_data, const int32_t* csr_indices,
                      const int32_t* csr_indptr, int64_t n, int64_t p, const float* v,
                      const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                      float* out, int accumulate, tm_stream_t stream);
int tm_csr_matvec_f64(const double* csr_data, const int32_t* csr_indices,
                      const int32_t* csr_indptr, int64_t n, int64_t p, const double* v,
                      const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                      double* out, int accumulate, tm_stream_t stream);

/* csc_rmatvec_unrestricted / csc_rmatvec, sparse.pyx:142-199.
 * res[c] = sum_{i in rows} X[i, cols[c]] * v[i].  accumulate: out[c] (+)= res[c]. */
int tm_csc_rmatvec_f32(const float* csc_data, const int32_t* csc_indices,
                       const int32_t* csc_indptr, int64_t n, int64_t p, const float* v,
                       const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                       float* out, int accumulate, tm_stream_t stream);
int tm_csc_rmatvec_f64(const double* csc_data, const int32_t* csc_indices,
                       const int32_t* csc_indptr, int64_t n, int64_t p, const double* v,
                       const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                       double* out, int accumulate, tm_stream_t stream);

/* csr_dense_sandwich (sparse.pyx:211-260), unrestricted, gather form: out = A^T diag(d) B for
 * a ROW-MAJOR B (n x q, q a multiple of 4 (f32) / 2 (f64), q <= 256 / 128) from a row-blocked CSC
 * copy of A: non-zeros ordered by (row block, column, row), bptr = n_blocks * p_sparse + 1
 * offsets, brow = global row ids.  Blocks should be small enough that block_rows * q * sizeof F
 * stays in L2 (<= 32 MB).  One vector RED per (row block, column) run instead of one per
 * non-zero; a `rows` restriction is expressed through d (zero weight outside).
 * Overwrites out (p_sparse x q). */
int tm_csc_dense_gather_sandwich_f32(const float* bdata, const int32_t* brow, const int32_t* bptr,
                                     int64_t p_sparse, int64_t n_blocks, const float* B, int64_t q,
                                     const float* d, float* out, tm_stream_t stream);
int tm_csc_dense_gather_sandwich_f64(const double* bdata, const int32_t* brow,
                                     const int32_t* bptr, int64_t p_sparse, int64_t n_blocks,
                                     const double* B, int64_t q, const double* d, double* out,
                                     tm_stream_t stream);

/* transpose_square_dot_weights (sparse), sparse.pyx:262-282.
 * out[j] = sum_{nz (i,j)} w[i] * X[i,j]^2.  Overwrites out (p). */
int tm_csc_sq_dot_weights_f32(const float* csc_data, const int32_t* csc_indices,
                              const int32_t* csc_indptr, int64_t n, int64_t p, const float* w,
                              float* out, tm_stream_t stream);
int tm_csc_sq_dot_weights_f64(const double* csc_data, const int32_t* csc_indices,
                              const int32_t* csc_indptr, int64_t n, int64_t p, const double* w,
                              double* out, tm_stream_t stream);

/* ---- categorical block (reference: ext/categorical.pyx, ext/split.pyx) -------------- */
/* `codes` is the int32 category index per row (read-only; -1 = missing);
 * `n_cat_cols` = number of matrix columns = #categories - drop_first; column of row k is
 * codes[k] - drop_first, rows with a negative column contribute nothing. */

/* sandwich_categorical_{fast,complex}, categorical.pyx:183-218.
 * out[codes[k]-drop_first] += d[k] for k in rows.  Overwrites out (n_cat_cols). */
int tm_cat_sandwich_f32(const int32_t* codes, int64_t n, const float* d, const int32_t* rows,
                        int64_t n_rows, int64_t n_cat_cols, int drop_first, float* out,
                        tm_stream_t stream);
int tm_cat_sandwich_f64(const int32_t* codes, int64_t n, const double* d, const int32_t* rows,
                        int64_t n_rows, int64_t n_cat_cols, int drop_first, double* out,
                        tm_stream_t stream);

/* transpose_matvec_{fast,complex}, categorical.pyx:23-117 ->
 * _transpose_matvec_all_rows_*, cat_split_helpers-tmpl.cpp:4-41.
 * out[c] += v[k] for k in rows, c = codes[k]-drop_first, c in cols.  ACCUMULATES into out
 * (length n_cat_cols) at the ABSOLUTE column index, like the reference. */
int tm_cat_transpose_matvec_f32(const int32_t* codes, int64_t n, const float* v,
                                const int32_t* rows, int64_t n_rows, const int32_t* cols,
                                int64_t n_cols, int64_t n_cat_cols, int drop_first, float* out,
                                tm_stream_t stream);
int tm_cat_transpose_matvec_f64(const int32_t* codes, int64_t n, const double* v,
                                const int32_t* rows, int64_t n_rows, const int32_t* cols,
                                int64_t n_cols, int64_t n_cat_cols, int drop_first, double* out,
                                tm_stream_t stream);

/* Deterministic form of the categorical histogram (transpose_matvec_fast/_complex,
 * categorical.pyx:23-126, and sandwich_categorical, :183-218): out[c] (+)= sum over the rows of
 * category c of w[k] (* row_w[k] when row_w != NULL), added in a fixed order, so the result is
 * bit-identical from run to run (the reference made this operation deterministic on purpose,
 * CHANGELOG.rst:134; the default kernels here add with atomics in arrival order).  perm = the
 * rows with a valid category ordered by category (stable), segptr[K + 1] = the category
 * offsets into perm.  accumulate = 0 overwrites out. */
int tm_cat_segment_sum_f32(const float* w, const float* row_w, const int32_t* perm,
                           const int32_t* segptr, int64_t K, float* out, int accumulate,
                           tm_stream_t stream);
int tm_cat_segment_sum_f64(const double* w, const double* row_w, const int32_t* perm,
                           const int32_t* segptr, int64_t K, double* out, int accumulate,
                           tm_stream_t stream);

/* multiply_complex / subset_categorical_complex, categorical.pyx:221-315: the CSR form of
 * diag(d) @ X for a categorical block, built on the device.  Row i owns one entry (column
 * codes[i] - drop_first, value d[i]) when codes[i] >= drop_first, none otherwise.
 * indptr[n + 1] is always written; indices / data (capacity n; data may be NULL = structure
 * only, d may be NULL = ones) receive the nnz = indptr[n] compacted entries. */
int tm_cat_to_csr_f32(const int32_t* codes, int64_t n, int drop_first, const float* d,
                      float* data, int32_t* indices, int32_t* indptr, tm_stream_t stream);
int tm_cat_to_csr_f64(const int32_t* codes, int64_t n, int drop_first, const double* d,
                      double* data, int32_t* indices, int32_t* indptr, tm_stream_t stream);

/* matvec_{fast,complex}, categorical.pyx:128-180.
 * out[i] += v[c], c = codes[i]-drop_first, c >= 0 and c in cols.  ACCUMULATES (length n). */
int tm_cat_matvec_f32(const int32_t* codes, int64_t n, const float* v, const int32_t* cols,
                      int64_t n_cols, int64_t n_cat_cols, int drop_first, float* out,
                      tm_stream_t stream);
int tm_cat_matvec_f64(const int32_t* codes, int64_t n, const double* v, const int32_t* cols,
                      int64_t n_cols, int64_t n_cat_cols, int drop_first, double* out,
                      tm_stream_t stream);

/* sandwich_cat_dense, split.pyx:32-80 -> _sandwich_cat_dense{C,F}_{fast,complex},
 * cat_split_helpers-tmpl.cpp:97-151.
 * out[(codes[k]-drop_first)*nJ + b] += d[k] * Y[k, j_cols[b]] for k in rows.
 * Overwrites out (n_cat_cols x nJ). */
int tm_cat_dense_sandwich_f32(const int32_t* codes, int64_t n, int64_t n_cat_cols,
                              int drop_first, const float* d, const float* Y, int64_t q,
                              int y_c_order, const int32_t* rows, int64_t n_rows,
                              const int32_t* j_cols, int64_t nJ, float* out,
                              tm_stream_t stream);
int tm_cat_dense_sandwich_f64(const int32_t* codes, int64_t n, int64_t n_cat_cols,
                              int drop_first, const double* d, const double* Y, int64_t q,
                              int y_c_order, const int32_t* rows, int64_t n_rows,
                              const int32_t* j_cols, int64_t nJ, double* out,
                              tm_stream_t stream);

/* sandwich_cat_cat, split.pyx:83-111 -> _sandwich_cat_cat_{fast,complex},
 * cat_split_helpers-tmpl.cpp:44-94.
 * out[(ci[k]-dfi)*Kj + (cj[k]-dfj)] += d[k] for k in rows.  Overwrites out (Ki x Kj). */
int tm_cat_cat_sandwich_f32(const int32_t* i_codes, const int32_t* j_codes, int64_t n,
                            int64_t Ki, int64_t Kj, int i_drop_first, int j_drop_first,
                            const float* d, const int32_t* rows, int64_t n_rows, float* out,
                            tm_stream_t stream);
int tm_cat_cat_sandwich_f64(const int32_t* i_codes, const int32_t* j_codes, int64_t n,
                            int64_t Ki, int64_t Kj, int i_drop_first, int j_drop_first,
                            const double* d, const int32_t* rows, int64_t n_rows, double* out,
                            tm_stream_t stream);

/* CategoricalMatrix._cross_sparse, categorical_matrix.py:825-838 (the reference runs this
 * block through scipy's csr_matmat; there is no reference native function).
 * out[(codes[k]-drop_first)*nS + s] += d[k] * A[k, s_cols[s]] for k in rows.
 * Overwrites out (n_cat_cols x nS). */
int tm_cat_sparse_sandwich_f32(const int32_t* codes, int64_t n, int64_t n_cat_cols,
                               int drop_first, const float* d, const float* csr_data,
                               const int32_t* csr_indices, const int32_t* csr_indptr,
                               const int32_t* csr_row, int64_t p_sparse, int64_t nnz,
                               const int32_t* rows, int64_t n_rows, const int32_t* s_cols,
                               int64_t nS, float* out, tm_stream_t stream);
int tm_cat_sparse_sandwich_f64(const int32_t* codes, int64_t n, int64_t n_cat_cols,
                               int drop_first, const double* d, const double* csr_data,
                               const int32_t* csr_indices, const int32_t* csr_indptr,
                               const int32_t* csr_row, int64_t p_sparse, int64_t nnz,
                               const int32_t* rows, int64_t n_rows, const int32_t* s_cols,
                               int64_t nS, double* out, tm_stream_t stream);

/* ---- fused cross blocks of a SplitMatrix (reference: the Python block loop,
 * split_matrix.py:346-354, which calls categorical_matrix.py:759-791 -> split.pyx:32-80 once per
 * categorical block and sparse_matrix.py:206-229 -> sparse.pyx:211-260 for the sparse block,
 * re-reading the dense block each time) ------------------------------------------------------ */
/* One pass over the C-contiguous dense block X (n x p, p % 4 == 0 for f32 / p % 2 == 0 for
 * f64, p <= 256 / 128):
 *   out_cat[i][(codes[i][k]-drop_first[i])*p + b] += d[k] * X[k, b]        i < n_cat (<= 8)
 *   out_sparse[j*p + b]                           += A[k, j] * d[k] * X[k, b]   (A in CSR)
 * for k in rows.  `codes`, `K`, `drop_first`, `out_cat` are HOST arrays of length n_cat (of
 * device pointers / values).  out_sparse may be NULL (no sparse block).  Overwrites outputs. */
int tm_dense_cross_sandwich_f32(const float* X, int64_t n, int64_t p, const float* d,
                                const int32_t* rows, int64_t n_rows, int n_cat,
                                const int32_t* const* codes, const int64_t* K,
                                const int32_t* drop_first, float* const* out_cat,
                                const float* csr_data, const int32_t* csr_indices,
                                const int32_t* csr_indptr, int64_t p_sparse, float* out_sparse,
                                tm_stream_t stream);
int tm_dense_cross_sandwich_f64(const double* X, int64_t n, int64_t p, const double* d,
                                const int32_t* rows, int64_t n_rows, int n_cat,
                                const int32_t* const* codes, const int64_t* K,
                                const int32_t* drop_first, double* const* out_cat,
                                const double* csr_data, const int32_t* csr_indices,
                                const int32_t* csr_indptr, int64_t p_sparse, double* out_sparse,
                                tm_stream_t stream);

/* Dense self block AND the dense x categorical cross blocks of categoricals with few levels in
 * one pass on the tensor cores (fp32, row-major X, p % 4 == 0, p <= 128, sum of K <= 320):
 * the weighted SYRK of tm_dense_sandwich_f32 plus one-hot MMAs
 *   out_cat[(off_i + codes[i][k] - drop_first[i]) * p + b] += d[k] * X[k, b],  off_i = sum_{c<i} K[c]
 * (reference: dense.pyx:19-44 + split.pyx:32-80 called per pair from split_matrix.py:337-354).
 * `codes`, `K`, `drop_first` are HOST arrays of length n_cat (<= 8).  out_dense (p x p) and
 * out_cat (sum K x p) are overwritten.  Inputs are rounded to TF32 (10-bit mantissa), sums are
 * accumulated in fp32. */
int tm_dense_onehot_sandwich_f32(const float* X, int64_t n, int64_t p, const float* d,
                                 const int32_t* rows, int64_t n_rows, int n_cat,
                                 const int32_t* const* codes, const int64_t* K,
                                 const int32_t* drop_first, float* out_dense, float* out_cat,
                                 tm_stream_t stream);

/* ---- SplitMatrix.sandwich as one native call (reference: split_matrix.py:324-356) ---------- */
/* One column block of a SplitMatrix (device pointers; `col_index` = positions of the block's
 * columns in the p x p result, split_matrix.py:232-247 `indices`). */
typedef struct tm_block_desc {
    int32_t kind;       /* 0 dense, 1 sparse (CSR + row ids), 2 categorical */
    int32_t c_order;    /* dense: 1 row-major, 0 column-major */
    int32_t drop_first; /* categorical */
    int32_t flags;      /* categorical: bit 0 = the rows are stored sorted by (among others) this
                         * block's codes, so equal codes form runs that the kernels pre-sum;
                         * bit 1 = this block is the primary sort key */
    int64_t ncols;      /* block width (categorical: #categories - drop_first) */
    const void* data;   /* dense: X (n x ncols); sparse: CSR data; categorical: int32 codes */
    const int32_t* csr_indices;
    const int32_t* csr_indptr;
    const int32_t* csr_row;
    int64_t nnz;
    const int64_t* col_index;
    /* categorical, optional (NULL = absent): rows with a valid category ordered by category
     * (stable) and the K+1 segment offsets into that list; enables the sorted-gather kernel
     * for the categorical x dense block */
    const int32_t* cat_perm;
    const int32_t* cat_segptr;
    int64_t cat_nvalid;
    /* sparse, optional (NULL = absent): the CSC copy of the block (values, row ids sorted inside
     * a column, ncols+1 offsets; sparse_matrix.py:133-143 keeps both forms too); enables the
     * atomics-free categorical x sparse kernel */
    const void* csc_data;
    const int32_t* csc_indices;
    const int32_t* csc_indptr;
    /* 0 / 1: plain CSC (ncols + 1 offsets).  B > 1: the CSC arrays are ROW-BLOCKED: the rows are
     * cut into B blocks of TM_CSC_ROW_BLOCK rows and the non-zeros are ordered by (row block,
     * column, row); csc_indptr then has B * ncols + 1 offsets, entry b * ncols + j = start of
     * column j inside row block b; csc_indices still holds global row ids.  Keeps the per-row
     * gathers of the categorical x sparse kernel inside the L2. */
    int64_t csc_row_blocks;
    /* sparse, optional: for every non-zero of the CSC arrays above (same order) the categorical
     * codes of its row, bit-packed: the categorical blocks of the SplitMatrix in block order,
     * block c in the next w_c = bit_length(ncols_c) bits from bit 0 upwards, value
     * code - drop_first, all ones = no category (missing / dropped); at most 64 bits in total.
     * Built once per matrix (like the reference's cached CSR, sparse_matrix.py:133-143): the
     * categorical x sparse kernel then streams the codes and gathers only d[row]. */
    const uint64_t* csc_cat_codes;
    /* sparse, optional: a second row-blocked CSC copy (same layout as csc_* with
     * csc_row_blocks > 1) whose blocks are small enough that a block of the DENSE operand stays
     * in L2 (block_rows * ncols_dense * sizeof F <= 32 MB): enables the gather form of the
     * dense x sparse block (one vector RED per (row block, column) run instead of one per
     * non-zero).  gcsc_row_blocks = number of row blocks B; gcsc_indptr has B * ncols + 1
     * offsets. */
    const void* gcsc_data;
    const int32_t* gcsc_indices;
    const int32_t* gcsc_indptr;
    int64_t gcsc_row_blocks;
} tm_block_desc;

#define TM_CSC_ROW_BLOCK (1 << 20)

/* sizeof(tm_block_desc) as compiled into the library (bindings check their struct mirror). */
int64_t tm_sizeof_block_desc(void);

/* Pass-level device timing of tm_split_sandwich_blocks_* (measurement aid): when enabled, CUDA
 * events bracket the three passes on the stream each one runs on; `ms[3]` = {tensor-core pass
 * (dense self + few-level categoricals, side stream), scatter pass (dense x many-level
 * categoricals and dense x sparse, caller's stream), index pass (all blocks without the dense
 * operand)}, averaged over the (at most 64 last) calls since tm_split_profile_enable(1); -1 for
 * a pass that did not run. */
/* Which forms the last tm_split_sandwich_blocks_* call of this process used (measurement aid):
 * bit 0 = the tcgen05 pass ran, bit 1 = the many-level categorical blocks were summed by its
 * scatter warps, bit 2 = so was dense x sparse, bit 3 = dense x sparse by row-blocked gather. */
int tm_split_last_plan(void);
void tm_split_profile_enable(int on);
int tm_split_profile_read(float* ms);

/* Elements (of the block dtype) of the flat workspace that holds every self block and every
 * cross block: for i: self_i (dense/sparse ncols_i^2, categorical ncols_i = the diagonal), then
 * for j > i: cross_ij (ncols_i * ncols_j); every block starts at a multiple of 4 elements. */
int64_t tm_split_workspace_elems(const tm_block_desc* blocks, int n_blocks);
/* Elements at the head of the workspace that belong to block 0 (its self block and its cross
 * blocks with every other block): when block 0 is the dense block this is exactly the part
 * tm_split_sandwich_blocks_part(…, 2) writes, the rest is what part 1 writes. */
int64_t tm_split_workspace_head_elems(const tm_block_desc* blocks, int n_blocks);
/* Every block of X[rows,:]^T diag(d[rows]) X[rows,:] into `workspace` (overwrites).  All blocks
 * share the dtype of the entry point; dense and sparse blocks must already be merged
 * (split_matrix.py:85-141).  Cross blocks that share the dense operand are computed in one
 * pass (tm_dense_cross_sandwich) when the dense block is row-major. */
int tm_split_sandwich_blocks_f32(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                 const float* d, const int32_t* rows, int64_t n_rows,
                                 float* workspace, tm_stream_t stream);
int tm_split_sandwich_blocks_f64(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                 const double* d, const int32_t* rows, int64_t n_rows,
                                 double* workspace, tm_stream_t stream);
/* Fused IRLS pass (callers of matrix_base.py:15-77 make two calls, sandwich(d) for the Hessian
 * and transpose_matvec(v) for the score, i.e. two passes over X): as tm_split_sandwich_blocks_*,
 * and in the same pass over the dense block dense_vec[c] = sum_k v[k] * X_dense[k, c] over `rows`
 * (ncols of the dense block; fp32/fp64 FMAs, never TF32).  The dense block is the only one whose
 * transpose_matvec costs a pass over ~all of the matrix bytes; the sparse and categorical parts
 * of X^T v are tm_csc_rmatvec / tm_cat_transpose_matvec calls.  No dense block: dense_vec is
 * not touched. */
int tm_split_sandwich_rmatvec_blocks_f32(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                         const float* d, const float* v, const int32_t* rows,
                                         int64_t n_rows, float* workspace, float* dense_vec,
                                         tm_stream_t stream);
int tm_split_sandwich_rmatvec_blocks_f64(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                         const double* d, const double* v, const int32_t* rows,
                                         int64_t n_rows, double* workspace, double* dense_vec,
                                         tm_stream_t stream);
/* Place the workspace into the p x p float64 result (split_matrix.py:336-354); `ld` = p.
 * Separate from the block computation so that a row-sharded caller can allreduce the flat
 * workspace in between.  A column whose entry in its block's `col_index` is negative is
 * dropped: with col_index[j] = position of column j in a sorted `cols` selection (or -1) and
 * ld = len(cols) this yields X[:, cols]^T D X[:, cols] (split.pyx:157-209 split_col_subsets +
 * split_matrix.py:334-354) from the same fused block computation. */
int tm_split_sandwich_assemble_f32(const tm_block_desc* blocks, int n_blocks,
                                   const float* workspace, double* out, int64_t ld,
                                   tm_stream_t stream);
int tm_split_sandwich_assemble_f64(const tm_block_desc* blocks, int n_blocks,
                                   const double* workspace, double* out, int64_t ld,
                                   tm_stream_t stream);

/* Rows [row0, row1) of the result only, written to out_band (row1 - row0 rows of `ld` values):
 * in a row-sharded job every rank places its own band after the allreduce of the workspace and
 * copies it to the host over its own PCIe link (N links instead of one). */
int tm_split_sandwich_assemble_band_f32(const tm_block_desc* blocks, int n_blocks,
                                        const float* workspace, double* out_band, int64_t ld,
                                        int64_t row0, int64_t row1, tm_stream_t stream);
int tm_split_sandwich_assemble_band_f64(const tm_block_desc* blocks, int n_blocks,
                                        const double* workspace, double* out_band, int64_t ld,
                                        int64_t row0, int64_t row1, tm_stream_t stream);

/* The same restricted to the blocks of `part` (1 = without a dense operand, 2 = with one; see
 * the two-phase calls below): a rank of a row-sharded job overlaps the host copy of its band. */
int tm_split_sandwich_assemble_part_band_f32(const tm_block_desc* blocks, int n_blocks,
                                             const float* workspace, double* out_band, int64_t ld,
                                             int part, int64_t row0, int64_t row1,
                                             tm_stream_t stream);
int tm_split_sandwich_assemble_part_band_f64(const tm_block_desc* blocks, int n_blocks,
                                             const double* workspace, double* out_band, int64_t ld,
                                             int part, int64_t row0, int64_t row1,
                                             tm_stream_t stream);

/* Two-phase form of the two calls above, for callers that want the result in HOST memory:
 * `part` 1 = the blocks without a dense operand (categorical / sparse self and cross blocks),
 * 2 = the blocks with one, 0 = all.  After blocks_part(1) + assemble_part(1) the finished
 * region of the p x p result can be copied to the host (tm_memcpy2d_to_host on another stream)
 * while blocks_part(2) still runs; the region is 96 % of the matrix at the benchmark shape. */
int tm_split_sandwich_blocks_part_f32(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                      const float* d, const int32_t* rows, int64_t n_rows,
                                      float* workspace, int part, tm_stream_t stream);
int tm_split_sandwich_blocks_part_f64(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                      const double* d, const int32_t* rows, int64_t n_rows,
                                      double* workspace, int part, tm_stream_t stream);
int tm_split_sandwich_assemble_part_f32(const tm_block_desc* blocks, int n_blocks,
                                        const float* workspace, double* out, int64_t ld, int part,
                                        tm_stream_t stream);
int tm_split_sandwich_assemble_part_f64(const tm_block_desc* blocks, int n_blocks,
                                        const double* workspace, double* out, int64_t ld, int part,
                                        tm_stream_t stream);
/* cudaMemcpy2DAsync device -> (pinned) host: `height` rows of `width_bytes`, pitches in bytes. */
int tm_memcpy2d_to_host(void* dst_host, int64_t dst_pitch, const void* src_dev, int64_t src_pitch,
                        int64_t width_bytes, int64_t height, tm_stream_t stream);

/* ---- row-order permutation (no reference counterpart: the reference keeps the caller's row
 * order; tabmat_b200.RowSortedMatrix stores the rows sorted by the many-level categorical codes
 * and maps the length-n vectors of sandwich / transpose_matvec / matvec through `perm`,
 * stored row i = original row perm[i]) ------------------------------------------------- */
/* dst[i] (+)= src[perm[i]], i < n. */
int tm_permute_gather_f32(const float* src, const int32_t* perm, int64_t n, float* dst,
                          int accumulate, tm_stream_t stream);
int tm_permute_gather_f64(const double* src, const int32_t* perm, int64_t n, double* dst,
                          int accumulate, tm_stream_t stream);
/* dst[perm[i]] (+)= src[i], i < n (perm must be a permutation: no duplicate targets). */
int tm_permute_scatter_f32(const float* src, const int32_t* perm, int64_t n, float* dst,
                           int accumulate, tm_stream_t stream);
int tm_permute_scatter_f64(const double* src, const int32_t* perm, int64_t n, double* dst,
                           int accumulate, tm_stream_t stream);

/* ---- StandardizedMatrix.sandwich epilogue (reference: standardized_mat.py:123-172) ---- */
/* out[i,j] = term1[i,j] * mult[i] * mult[j] + dm[i] * shift[j] + shift[i] * dm[j]
 *            + shift[i] * shift[j] * sum_d,   dm[i] = d_mat[i] * mult[i]   (m x m, overwrites).
 * term1 = the inner matrix' sandwich restricted like the call: float64 when term1_f64 != 0 (a
 * SplitMatrix), else the entry point's dtype; `diag` != 0: term1 is a categorical block's
 * diagonal (m values).  d_mat = inner.transpose_matvec(d); mult may be NULL (no scaling);
 * sum_d points to a DEVICE scalar (sum of d over the rows), so nothing synchronises. */
int tm_std_sandwich_combine_f32(const void* term1, int term1_f64, int diag, const float* dmat,
                                const float* shift, const float* mult, const float* sum_d,
                                int64_t m, float* out, tm_stream_t stream);
int tm_std_sandwich_combine_f64(const void* term1, int term1_f64, int diag, const double* dmat,
                                const double* shift, const double* mult, const double* sum_d,
                                int64_t m, double* out, tm_stream_t stream);

/* ---- SplitMatrix assembly (reference: split_matrix.py:336-354, the numpy scatter) ---- */
/* out[ri[a]*ld + ci[b]] = blk[a*nb + b]  (and, when mirror != 0, out[ci[b]*ld + ri[a]] too).
 * ri / ci NULL = identity.  `out` is float64 (SplitMatrix.sandwich always returns float64,
 * split_matrix.py:336). */
int tm_scatter_block_f32(const float* blk, int64_t na, int64_t nb, const int64_t* ri,
                         const int64_t* ci, double* out, int64_t ld, int mirror,
                         tm_stream_t stream);
int tm_scatter_block_f64(const double* blk, int64_t na, int64_t nb, const int64_t* ri,
                         const int64_t* ci, double* out, int64_t ld, int mirror,
                         tm_stream_t stream);
/* Categorical self block: out[ri[a]*ld + ri[b]] = (a==b) ? diag[a] : 0. */
int tm_scatter_diag_f32(const float* diag, int64_t na, const int64_t* ri, double* out,
                        int64_t ld, tm_stream_t stream);
int tm_scatter_diag_f64(const double* diag, int64_t na, const int64_t* ri, double* out,
                        int64_t ld, tm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TABMAT_B200_H */
