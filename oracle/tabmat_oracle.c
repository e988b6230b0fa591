/*
 * tabmat_oracle.c — CPU restatement of the reference's hot-path algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tabmat_b200/ links, imports or executes this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the
 * checker.  Parity status: PINNED — checked in tests/test_oracle_cpu.py against (a) dense
 * numpy recomputation (the reference's own test strategy), (b) the golden fixtures under
 * tests/golden/ generated from the reference package itself, and (c) the reference's
 * compiled kernels in oracle/_ref when present.
 *
 * Every function is a plain, single-threaded, double-loop restatement; the reference
 * file:line it follows is cited above it (paths relative to the reference tree,
 * src/tabmat/ext/...).  Loop ORDER is not reproduced (the reference blocks and
 * parallelises); the arithmetic per output element is.
 *
 * Built by oracle/build_oracle.py with plain gcc; instantiated for float and double.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ROW(rows, t) ((rows) ? (int64_t)(rows)[t] : (int64_t)(t))
#define COL(cols, c) ((cols) ? (int64_t)(cols)[c] : (int64_t)(c))
#define XAT(X, k, j, n, p, c_order) ((c_order) ? (X)[(k) * (p) + (j)] : (X)[(j) * (n) + (k)])

#define DEFINE_ORACLE(F, SUF)                                                                   \
                                                                                                \
    /* dense.pyx:19-44 -> dense_helpers-tmpl.cpp:266-308 (lower triangle, then mirrored) */     \
    void orc_dense_sandwich_##SUF(const F* X, int64_t n, int64_t p, int c_order, const F* d,    \
                                  const int32_t* rows, int64_t n_rows, const int32_t* cols,     \
                                  int64_t m, F* out) {                                          \
        memset(out, 0, sizeof(F) * (size_t)(m * m));                                            \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            F dk = d[k];                                                                        \
            for (int64_t a = 0; a < m; ++a) {                                                   \
                F xa = XAT(X, k, COL(cols, a), n, p, c_order) * dk;                             \
                for (int64_t b = 0; b <= a; ++b)                                                \
                    out[a * m + b] += xa * XAT(X, k, COL(cols, b), n, p, c_order);              \
            }                                                                                   \
        }                                                                                       \
        for (int64_t a = 0; a < m; ++a)                                                         \
            for (int64_t b = 0; b < a; ++b) out[b * m + a] = out[a * m + b];                    \
    }                                                                                           \
                                                                                                \
    /* dense.pyx:76-101 -> dense_helpers-tmpl.cpp:385-417 */                                    \
    void orc_dense_matvec_##SUF(const F* X, int64_t n, int64_t p, int c_order, const F* v,      \
                                const int32_t* rows, int64_t n_rows, const int32_t* cols,       \
                                int64_t n_cols, F* out) {                                       \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            F s = 0;                                                                            \
            for (int64_t c = 0; c < n_cols; ++c) {                                              \
                int64_t j = COL(cols, c);                                                       \
                s += XAT(X, k, j, n, p, c_order) * v[j];                                        \
            }                                                                                   \
            out[t] = s;                                                                         \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* dense.pyx:48-73 -> dense_helpers-tmpl.cpp:314-383 */                                     \
    void orc_dense_rmatvec_##SUF(const F* X, int64_t n, int64_t p, int c_order, const F* v,     \
                                 const int32_t* rows, int64_t n_rows, const int32_t* cols,      \
                                 int64_t n_cols, F* out) {                                      \
        for (int64_t c = 0; c < n_cols; ++c) out[c] = 0;                                        \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            for (int64_t c = 0; c < n_cols; ++c)                                                \
                out[c] += XAT(X, k, COL(cols, c), n, p, c_order) * v[k];                        \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* dense.pyx:103-122 */                                                                     \
    void orc_dense_sq_dot_weights_##SUF(const F* X, int64_t n, int64_t p, int c_order,          \
                                        const F* w, const F* shift, F* out) {                   \
        for (int64_t j = 0; j < p; ++j) {                                                       \
            F s = 0;                                                                            \
            for (int64_t i = 0; i < n; ++i) {                                                   \
                F t = XAT(X, i, j, n, p, c_order) - shift[j];                                   \
                s += w[i] * t * t;                                                              \
            }                                                                                   \
            out[j] = s;                                                                         \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* sparse.pyx:17-77: CSC outer loop over j, CSR walk of row k while i <= j, row mask,    */ \
    /* col_map, tril mirror                                                                  */ \
    void orc_sparse_sandwich_##SUF(const F* csc_data, const int32_t* csc_indices,               \
                                   const int32_t* csc_indptr, const F* csr_data,                \
                                   const int32_t* csr_indices, const int32_t* csr_indptr,       \
                                   int64_t n, int64_t p, const F* d, const int32_t* rows,       \
                                   int64_t n_rows, const int32_t* cols, int64_t m, F* out) {    \
        memset(out, 0, sizeof(F) * (size_t)(m * m));                                            \
        uint8_t* row_included = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);                   \
        int32_t* col_map = (int32_t*)malloc(sizeof(int32_t) * (size_t)(p > 0 ? p : 1));         \
        for (int64_t t = 0; t < n_rows; ++t) row_included[ROW(rows, t)] = 1;                    \
        for (int64_t j = 0; j < p; ++j) col_map[j] = -1;                                        \
        for (int64_t c = 0; c < m; ++c) col_map[COL(cols, c)] = (int32_t)c;                     \
        for (int64_t Cj = 0; Cj < m; ++Cj) {                                                    \
            int64_t j = COL(cols, Cj);                                                          \
            for (int64_t a = csc_indptr[j]; a < csc_indptr[j + 1]; ++a) {                       \
                int64_t k = csc_indices[a];                                                     \
                if (!row_included[k]) continue;                                                 \
                F A_val = csc_data[a] * d[k];                                                   \
                for (int64_t b = csr_indptr[k]; b < csr_indptr[k + 1]; ++b) {                   \
                    int64_t i = csr_indices[b];                                                 \
                    if (i > j) break;                                                           \
                    int32_t Ci = col_map[i];                                                    \
                    if (Ci == -1) continue;                                                     \
                    out[Cj * m + Ci] += csr_data[b] * A_val;                                    \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
        for (int64_t a = 0; a < m; ++a)                                                         \
            for (int64_t b = 0; b < a; ++b) out[b * m + a] += out[a * m + b];                   \
        free(row_included);                                                                     \
        free(col_map);                                                                          \
    }                                                                                           \
                                                                                                \
    /* sparse.pyx:211-260 -> sparse_helpers-tmpl.cpp:23-143 */                                  \
    void orc_csr_dense_sandwich_##SUF(const F* data, const int32_t* indices,                    \
                                      const int32_t* indptr, int64_t n, int64_t p_sparse,       \
                                      const F* B, int64_t q, int b_c_order, const F* d,         \
                                      const int32_t* rows, int64_t n_rows,                      \
                                      const int32_t* a_cols, int64_t nA, const int32_t* b_cols, \
                                      int64_t nB, F* out) {                                     \
        memset(out, 0, sizeof(F) * (size_t)(nA * nB));                                          \
        int32_t* col_map = (int32_t*)malloc(sizeof(int32_t) * (size_t)(p_sparse > 0 ? p_sparse : 1)); \
        for (int64_t j = 0; j < p_sparse; ++j) col_map[j] = -1;                                 \
        for (int64_t c = 0; c < nA; ++c) col_map[COL(a_cols, c)] = (int32_t)c;                  \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            for (int64_t e = indptr[k]; e < indptr[k + 1]; ++e) {                               \
                int32_t Ci = col_map[indices[e]];                                               \
                if (Ci == -1) continue;                                                         \
                F Q = data[e];                                                                  \
                for (int64_t b = 0; b < nB; ++b)                                                \
                    out[(int64_t)Ci * nB + b] +=                                                \
                        Q * (d[k] * XAT(B, k, COL(b_cols, b), n, q, b_c_order));                \
            }                                                                                   \
        }                                                                                       \
        free(col_map);                                                                          \
    }                                                                                           \
                                                                                                \
    /* sparse.pyx:79-140 (restricted form; rows/cols NULL = unrestricted) */                    \
    void orc_csr_matvec_##SUF(const F* data, const int32_t* indices, const int32_t* indptr,     \
                              int64_t n, int64_t p, const F* v, const int32_t* rows,            \
                              int64_t n_rows, const int32_t* cols, int64_t n_cols, F* out) {    \
        uint8_t* inc = (uint8_t*)calloc((size_t)(p > 0 ? p : 1), 1);                            \
        for (int64_t c = 0; c < n_cols; ++c) inc[COL(cols, c)] = 1;                             \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            F s = 0;                                                                            \
            for (int64_t e = indptr[k]; e < indptr[k + 1]; ++e)                                 \
                if (inc[indices[e]]) s += data[e] * v[indices[e]];                              \
            out[t] = s;                                                                         \
        }                                                                                       \
        free(inc);                                                                              \
    }                                                                                           \
                                                                                                \
    /* sparse.pyx:142-199 */                                                                    \
    void orc_csc_rmatvec_##SUF(const F* data, const int32_t* indices, const int32_t* indptr,    \
                               int64_t n, int64_t p, const F* v, const int32_t* rows,           \
                               int64_t n_rows, const int32_t* cols, int64_t n_cols, F* out) {   \
        uint8_t* inc = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);                            \
        for (int64_t t = 0; t < n_rows; ++t) inc[ROW(rows, t)] = 1;                             \
        for (int64_t c = 0; c < n_cols; ++c) {                                                  \
            int64_t j = COL(cols, c);                                                           \
            F s = 0;                                                                            \
            for (int64_t e = indptr[j]; e < indptr[j + 1]; ++e)                                 \
                if (inc[indices[e]]) s += data[e] * v[indices[e]];                              \
            out[c] = s;                                                                         \
        }                                                                                       \
        free(inc);                                                                              \
    }                                                                                           \
                                                                                                \
    /* sparse.pyx:262-282 */                                                                    \
    void orc_csc_sq_dot_weights_##SUF(const F* data, const int32_t* indices,                    \
                                      const int32_t* indptr, int64_t p, const F* w, F* out) {   \
        for (int64_t j = 0; j < p; ++j) {                                                       \
            F s = 0;                                                                            \
            for (int64_t e = indptr[j]; e < indptr[j + 1]; ++e)                                 \
                s += w[indices[e]] * data[e] * data[e];                                         \
            out[j] = s;                                                                         \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* categorical.pyx:183-218 */                                                               \
    void orc_cat_sandwich_##SUF(const int32_t* codes, const F* d, const int32_t* rows,          \
                                int64_t n_rows, int64_t K, int drop_first, F* out) {            \
        for (int64_t c = 0; c < K; ++c) out[c] = 0;                                             \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            int64_t c = (int64_t)codes[k] - drop_first;                                         \
            if (c >= 0) out[c] += d[k];                                                         \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* categorical.pyx:23-117 -> cat_split_helpers-tmpl.cpp:4-41; adds at ABSOLUTE column */    \
    void orc_cat_transpose_matvec_##SUF(const int32_t* codes, const F* v, const int32_t* rows,  \
                                        int64_t n_rows, const int32_t* cols, int64_t n_cols,    \
                                        int64_t K, int drop_first, F* out) {                    \
        uint8_t* inc = (uint8_t*)calloc((size_t)(K > 0 ? K : 1), 1);                            \
        for (int64_t c = 0; c < (cols ? n_cols : K); ++c) inc[COL(cols, c)] = 1;                \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            int64_t c = (int64_t)codes[k] - drop_first;                                         \
            if (c >= 0 && inc[c]) out[c] += v[k];                                               \
        }                                                                                       \
        free(inc);                                                                              \
    }                                                                                           \
                                                                                                \
    /* categorical.pyx:128-180; accumulates */                                                  \
    void orc_cat_matvec_##SUF(const int32_t* codes, int64_t n, const F* v, const int32_t* cols, \
                              int64_t n_cols, int64_t K, int drop_first, F* out) {              \
        uint8_t* inc = (uint8_t*)calloc((size_t)(K > 0 ? K : 1), 1);                            \
        for (int64_t c = 0; c < (cols ? n_cols : K); ++c) inc[COL(cols, c)] = 1;                \
        for (int64_t i = 0; i < n; ++i) {                                                       \
            int64_t c = (int64_t)codes[i] - drop_first;                                         \
            if (c >= 0 && inc[c]) out[i] += v[c];                                               \
        }                                                                                       \
        free(inc);                                                                              \
    }                                                                                           \
                                                                                                \
    /* split.pyx:32-80 -> cat_split_helpers-tmpl.cpp:97-151 */                                  \
    void orc_cat_dense_sandwich_##SUF(const int32_t* codes, int64_t n, int64_t K,               \
                                      int drop_first, const F* d, const F* Y, int64_t q,        \
                                      int y_c_order, const int32_t* rows, int64_t n_rows,       \
                                      const int32_t* j_cols, int64_t nJ, F* out) {              \
        memset(out, 0, sizeof(F) * (size_t)(K * nJ));                                           \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            int64_t i = (int64_t)codes[k] - drop_first;                                         \
            if (i < 0) continue;                                                                \
            for (int64_t b = 0; b < nJ; ++b)                                                    \
                out[i * nJ + b] += d[k] * XAT(Y, k, COL(j_cols, b), n, q, y_c_order);           \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* split.pyx:83-111 -> cat_split_helpers-tmpl.cpp:44-94 */                                  \
    void orc_cat_cat_sandwich_##SUF(const int32_t* ic, const int32_t* jc, int64_t Ki,           \
                                    int64_t Kj, int dfi, int dfj, const F* d,                   \
                                    const int32_t* rows, int64_t n_rows, F* out) {              \
        memset(out, 0, sizeof(F) * (size_t)(Ki * Kj));                                          \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            int64_t i = (int64_t)ic[k] - dfi, j = (int64_t)jc[k] - dfj;                         \
            if (i < 0 || j < 0) continue;                                                       \
            out[i * Kj + j] += d[k];                                                            \
        }                                                                                       \
    }                                                                                           \
                                                                                                \
    /* categorical_matrix.py:825-838: (diag(d) X_cat)[rows, L]^T @ A[rows, R]; the reference */ \
    /* computes it with scipy csr_matmat; the arithmetic per output cell is restated here    */ \
    void orc_cat_sparse_sandwich_##SUF(const int32_t* codes, int64_t n, int64_t K,              \
                                       int drop_first, const F* d, const F* data,               \
                                       const int32_t* indices, const int32_t* indptr,           \
                                       int64_t p_sparse, const int32_t* rows, int64_t n_rows,   \
                                       const int32_t* s_cols, int64_t nS, F* out) {             \
        memset(out, 0, sizeof(F) * (size_t)(K * nS));                                           \
        int32_t* col_map = (int32_t*)malloc(sizeof(int32_t) * (size_t)(p_sparse > 0 ? p_sparse : 1)); \
        for (int64_t j = 0; j < p_sparse; ++j) col_map[j] = -1;                                 \
        for (int64_t c = 0; c < nS; ++c) col_map[COL(s_cols, c)] = (int32_t)c;                  \
        for (int64_t t = 0; t < n_rows; ++t) {                                                  \
            int64_t k = ROW(rows, t);                                                           \
            int64_t i = (int64_t)codes[k] - drop_first;                                         \
            if (i < 0) continue;                                                                \
            for (int64_t e = indptr[k]; e < indptr[k + 1]; ++e) {                               \
                int32_t s = col_map[indices[e]];                                                \
                if (s >= 0) out[i * nS + s] += d[k] * data[e];                                  \
            }                                                                                   \
        }                                                                                       \
        free(col_map);                                                                          \
    }

DEFINE_ORACLE(float, f32)
DEFINE_ORACLE(double, f64)
