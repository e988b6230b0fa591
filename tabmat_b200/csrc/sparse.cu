// Sparse (CSR/CSC) block kernels: self sandwich, sparse x dense cross sandwich,
// SpMV / SpMV^T with row/column restrictions, weighted column second moment.
//
// Reference semantics: sparse.pyx:17-282, sparse_helpers-tmpl.cpp:23-143.
#include "tm_common.cuh"

namespace tmb {

// ---------------------------------------------------------------------------------------
// Self sandwich, CSR row outer products.  One thread per non-zero e = (k, j_a):
//   for every non-zero b <= e of the same row:  out[pos(j_a), pos(j_b)] += v_a * d_k * v_b
// Column indices are sorted inside a row and `cols` is sorted, so pos(j_b) <= pos(j_a): only
// the lower triangle is touched (the reference does the same with its `i > j: break`,
// sparse.pyx:64-67) and mirrored afterwards (sparse.pyx:76).
// ---------------------------------------------------------------------------------------
// PACKED: `out` is the packed lower triangle (row pa starts at pa (pa + 1) / 2) - half the
// footprint, so that the table of a wide block stays in the 126 MB L2 instead of doing DRAM
// read-modify-writes (C4: 5000^2 doubles = 200 MB as a square, 100 MB packed); k_unpack_tri
// writes the square afterwards.
template <typename F, bool OFFDIAG = false, bool PACKED = false>
__global__ void k_sparse_sandwich(const F* __restrict__ data, const int32_t* __restrict__ indices,
                                  const int32_t* __restrict__ indptr,
                                  const int32_t* __restrict__ nz_row, int64_t nnz,
                                  const F* __restrict__ d, const uint8_t* __restrict__ row_mask,
                                  const int32_t* __restrict__ col_pos, int64_t m,
                                  F* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
        int k = nz_row[e];
        if (row_mask && !row_mask[k]) continue;
        int ja = indices[e];
        int pa = col_pos ? col_pos[ja] : ja;
        if (pa < 0) continue;
        F va = data[e] * d[k];
        F* orow = out + (PACKED ? (int64_t)pa * (pa + 1) / 2 : (int64_t)pa * m);
        for (int64_t b = indptr[k]; b < e + (OFFDIAG ? 0 : 1); ++b) {
            int jb = indices[b];
            int pb = col_pos ? col_pos[jb] : jb;
            if (pb < 0) continue;
            red_add(&orow[pb], va * data[b]);
        }
    }
}

// packed lower triangle -> full symmetric square (every element of `out` is written)
template <typename F>
__global__ void k_unpack_tri(const F* __restrict__ tri, int64_t m, F* __restrict__ out) {
    const int64_t total = m * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t i = t / m, j = t - i * m;
        const int64_t a = i > j ? i : j, b = i > j ? j : i;
        out[t] = tri[a * (a + 1) / 2 + b];
    }
}

// The same with the (packed lower-triangular) result table of a NARROW sparse block in shared
// memory: m (m + 1) / 2 entries <= ~200 KB (m <= 319 in f32, 225 in f64).  One CTA per SM keeps
// a private table, adds with shared-memory atomics (4x the L2 RED rate of the whole chip, no
// contention collapse on the few popular columns) and flushes its non-zero entries with one RED
// each.  Pays once nnz is well above SMs * table size (the flush); the host picks.
constexpr int SS_THREADS = 1024;
template <typename F>
__global__ void __launch_bounds__(SS_THREADS)
k_sparse_sandwich_smem(const F* __restrict__ data, const int32_t* __restrict__ indices,
                       const int32_t* __restrict__ indptr, const int32_t* __restrict__ nz_row,
                       int64_t nnz, const F* __restrict__ d, const uint8_t* __restrict__ row_mask,
                       const int32_t* __restrict__ col_pos, int m, F* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char ss_raw[];
    F* tab = reinterpret_cast<F*>(ss_raw);
    const int tri = m * (m + 1) / 2;
    for (int i = threadIdx.x; i < tri; i += SS_THREADS) tab[i] = F(0);
    __syncthreads();
    // a CTA takes contiguous pieces of the non-zeros: the inner walk over the row stays in L1
    const int64_t stride = (int64_t)gridDim.x * SS_THREADS;
    for (int64_t e = (int64_t)blockIdx.x * SS_THREADS + threadIdx.x; e < nnz; e += stride) {
        const int k = nz_row[e];
        if (row_mask && !row_mask[k]) continue;
        const int ja = indices[e];
        const int pa = col_pos ? col_pos[ja] : ja;
        if (pa < 0) continue;
        const F dk = d[k];
        if (dk == F(0)) continue;
        const F va = data[e] * dk;
        F* trow = tab + pa * (pa + 1) / 2;
        for (int b = indptr[k]; b <= e; ++b) {
            const int jb = indices[b];
            const int pb = col_pos ? col_pos[jb] : jb;
            if (pb < 0) continue;
            atomicAdd(&trow[pb], va * data[b]);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < m * m; idx += SS_THREADS) {
        const int pa = idx / m, pb = idx - pa * m;
        if (pb > pa) continue;
        const F v = tab[pa * (pa + 1) / 2 + pb];
        if (v != F(0)) red_add(&out[idx], v);
    }
}

// ---------------------------------------------------------------------------------------
// sparse x dense cross sandwich.  One warp per row k: the lanes hold d_k * B[k, B_cols[.]]
// in registers and, for every non-zero (k, j) with pos(j) >= 0, RED-add the scaled row into
// out[pos(j), :] (consecutive addresses across the warp).
// ---------------------------------------------------------------------------------------
template <typename F, bool C_ORDER, int MAXQ>
__global__ void __launch_bounds__(256)
k_csr_dense(const F* __restrict__ data, const int32_t* __restrict__ indices,
            const int32_t* __restrict__ indptr, const F* __restrict__ B, int64_t n, int64_t q,
            const F* __restrict__ d, const int32_t* __restrict__ rows, int64_t n_rows,
            const int32_t* __restrict__ a_pos, const int32_t* __restrict__ bcols, int64_t nB,
            int64_t b_off, F* __restrict__ out) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp; t < n_rows; t += nwarps) {
        int64_t k = row_at(rows, t);
        int e0 = indptr[k], e1 = indptr[k + 1];
        if (e0 == e1) continue;
        F dk = d[k];
        F r[MAXQ];
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) {
            int64_t b = b_off + lane + 32 * i;
            F y = F(0);
            if (b < nB) {
                int64_t j = bcols ? (int64_t)bcols[b] : b;
                y = C_ORDER ? B[k * q + j] : B[j * n + k];
            }
            r[i] = dk * y;
        }
        for (int e = e0; e < e1; ++e) {
            int ja = indices[e];
            int pa = a_pos ? a_pos[ja] : ja;
            if (pa < 0) continue;
            F va = data[e];
            F* orow = out + (int64_t)pa * nB;
#pragma unroll
            for (int i = 0; i < MAXQ; ++i) {
                int64_t b = b_off + lane + 32 * i;
                if (b < nB) red_add(&orow[b], va * r[i]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// CSR SpMV (thread per row):  out[t] (+)= sum_{j in cols} X[rows[t], j] * v[j]
// ---------------------------------------------------------------------------------------
template <typename F>
__global__ void k_csr_matvec(const F* __restrict__ data, const int32_t* __restrict__ indices,
                             const int32_t* __restrict__ indptr, const F* __restrict__ v,
                             const int32_t* __restrict__ rows, int64_t n_rows,
                             const uint8_t* __restrict__ col_mask, F* __restrict__ out,
                             int accumulate) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += stride) {
        int64_t k = row_at(rows, t);
        F s = F(0);
        for (int e = indptr[k]; e < indptr[k + 1]; ++e) {
            int j = indices[e];
            if (col_mask && !col_mask[j]) continue;
            s = fma(data[e], v[j], s);
        }
        out[t] = accumulate ? out[t] + s : s;
    }
}

// ---------------------------------------------------------------------------------------
// CSC column reductions (warp per column):
//   MODE 0: out[c] (+)= sum_{i in rows} X[i, cols[c]] * v[i]
//   MODE 1: out[c]   = sum_i w[i] * X[i, c]^2
// ---------------------------------------------------------------------------------------
template <typename F, int MODE>
__global__ void k_csc_colreduce(const F* __restrict__ data, const int32_t* __restrict__ indices,
                                const int32_t* __restrict__ indptr, const F* __restrict__ v,
                                const uint8_t* __restrict__ row_mask,
                                const int32_t* __restrict__ cols, int64_t n_cols,
                                F* __restrict__ out, int accumulate) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < n_cols; c += nwarps) {
        int64_t j = cols ? (int64_t)cols[c] : c;
        F s = F(0);
        const int e_end = indptr[j + 1];
        int e = indptr[j] + lane;
        // four index loads, then four dependent gathers of v in flight per lane
        for (; e + 96 < e_end; e += 128) {
            int i[4];
            F x[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                i[u] = __ldg(indices + e + 32 * u);
                x[u] = __ldg(data + e + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                w[u] = (row_mask && !row_mask[i[u]]) ? F(0) : __ldg(v + i[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) s += (MODE == 0) ? x[u] * w[u] : w[u] * x[u] * x[u];
        }
        for (; e < e_end; e += 32) {
            int i = indices[e];
            if (row_mask && !row_mask[i]) continue;
            F x = data[e];
            s += (MODE == 0) ? x * v[i] : v[i] * x * x;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[c] = accumulate ? out[c] + s : s;
    }
}

// ---- host wrappers ---------------------------------------------------------------------
template <typename F>
int sparse_sandwich_ex(const F* data, const int32_t* indices, const int32_t* indptr,
                       const int32_t* nz_row, int64_t n, int64_t p, int64_t nnz, const F* d,
                       const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t m, F* out,
                       cudaStream_t st, bool offdiag_only) {
    if (!cols) m = p;
    if (m <= 0) return 0;
    // a square that does not fit the L2 is accumulated as a packed triangle (TABMAT_B200_SPARSE_TRI:
    // 0 never, 1 when the square is > 96 MB (default), 2 always)
    const int tri_mode = getenv("TABMAT_B200_SPARSE_TRI") ? atoi(getenv("TABMAT_B200_SPARSE_TRI")) : 1;
    const size_t sq_bytes = sizeof(F) * (size_t)(m * m);
    const bool packed = !offdiag_only && nnz > 0 && !(rows && n_rows <= 0) &&
                        (tri_mode == 2 || (tri_mode == 1 && sq_bytes > ((size_t)96 << 20)));
    const size_t tri_elems = (size_t)(m * (m + 1) / 2);
    Scratch tri(packed ? sizeof(F) * tri_elems : 0, st);
    if (tri.err != cudaSuccess) return fail_cuda(tri.err, "scratch");
    if (packed) {
        TM_CUDA(cudaMemsetAsync(tri.p, 0, sizeof(F) * tri_elems, st));
    } else if (!offdiag_only) {
        TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)(m * m), st));
    }
    if (nnz <= 0 || (rows && n_rows <= 0)) return 0;
    Scratch rmask(rows ? (size_t)n : 0, st);
    Scratch cpos(cols ? sizeof(int32_t) * (size_t)p : 0, st);
    if (rmask.err != cudaSuccess) return fail_cuda(rmask.err, "scratch");
    if (cpos.err != cudaSuccess) return fail_cuda(cpos.err, "scratch");
    if (rows) {
        int rc = build_mask(rows, n_rows, n, rmask.as<uint8_t>(), st);
        if (rc) return rc;
    }
    if (cols) {
        int rc = build_pos_map(cols, m, p, cpos.as<int32_t>(), st);
        if (rc) return rc;
    }
    // narrow block, many non-zeros: private shared-memory tables (TABMAT_B200_SPARSE_SMEM=0
    // keeps the L2 RED form, =2 forces the tables whenever they fit)
    const size_t tri_bytes = sizeof(F) * (size_t)(m * (m + 1) / 2);
    const char* sm_env = getenv("TABMAT_B200_SPARSE_SMEM");
    const int sm_mode = sm_env ? atoi(sm_env) : 1;
    const bool in_smem = !offdiag_only && !packed && sm_mode != 0 && tri_bytes <= 200 * 1024 &&
                         (sm_mode == 2 || nnz > (int64_t)sm_count() * (m * (m + 1) / 2));
    if (in_smem) {
        static bool attr_done[2][16] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        bool& done = attr_done[sizeof(F) == 8][dev & 15];
        if (!done) {
            TM_CUDA(cudaFuncSetAttribute(k_sparse_sandwich_smem<F>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            done = true;
        }
        const int g = (int)std::min<int64_t>(sm_count(), (nnz + SS_THREADS - 1) / SS_THREADS);
        k_sparse_sandwich_smem<F><<<g, SS_THREADS, tri_bytes, st>>>(
            data, indices, indptr, nz_row, nnz, d, rows ? rmask.as<uint8_t>() : nullptr,
            cols ? cpos.as<int32_t>() : nullptr, (int)m, out);
    } else {
        int g = grid_for(nnz, 256, sm_count() * 32);
        if (packed) {
            k_sparse_sandwich<F, false, true><<<g, 256, 0, st>>>(
                data, indices, indptr, nz_row, nnz, d, rows ? rmask.as<uint8_t>() : nullptr,
                cols ? cpos.as<int32_t>() : nullptr, m, tri.as<F>());
            TM_LAUNCHED();
            k_unpack_tri<F><<<grid_for(m * m, 256, sm_count() * 16), 256, 0, st>>>(tri.as<F>(), m, out);
            TM_LAUNCHED();
            return 0;
        }
        if (offdiag_only)
            k_sparse_sandwich<F, true><<<g, 256, 0, st>>>(data, indices, indptr, nz_row, nnz, d,
                                                          rows ? rmask.as<uint8_t>() : nullptr,
                                                          cols ? cpos.as<int32_t>() : nullptr, m, out);
        else
            k_sparse_sandwich<F><<<g, 256, 0, st>>>(data, indices, indptr, nz_row, nnz, d,
                                                    rows ? rmask.as<uint8_t>() : nullptr,
                                                    cols ? cpos.as<int32_t>() : nullptr, m, out);
    }
    TM_LAUNCHED();
    return symmetrize_from_lower<F>(out, m, st);
}

template int sparse_sandwich_ex<float>(const float*, const int32_t*, const int32_t*, const int32_t*,
                                       int64_t, int64_t, int64_t, const float*, const int32_t*,
                                       int64_t, const int32_t*, int64_t, float*, cudaStream_t, bool);
template int sparse_sandwich_ex<double>(const double*, const int32_t*, const int32_t*,
                                        const int32_t*, int64_t, int64_t, int64_t, const double*,
                                        const int32_t*, int64_t, const int32_t*, int64_t, double*,
                                        cudaStream_t, bool);

template <typename F>
int csr_dense_sandwich(const F* data, const int32_t* indices, const int32_t* indptr, int64_t n,
                       int64_t p_sparse, const F* B, int64_t q, int b_c_order, const F* d,
                       const int32_t* rows, int64_t n_rows, const int32_t* a_cols, int64_t nA,
                       const int32_t* b_cols, int64_t nB, F* out, cudaStream_t st) {
    if (!a_cols) nA = p_sparse;
    if (!b_cols) nB = q;
    if (!rows) n_rows = n;
    if (nA <= 0 || nB <= 0) return 0;
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)(nA * nB), st));
    if (n_rows <= 0) return 0;
    Scratch apos(a_cols ? sizeof(int32_t) * (size_t)p_sparse : 0, st);
    if (apos.err != cudaSuccess) return fail_cuda(apos.err, "scratch");
    if (a_cols) {
        int rc = build_pos_map(a_cols, nA, p_sparse, apos.as<int32_t>(), st);
        if (rc) return rc;
    }
    const int32_t* ap = a_cols ? apos.as<int32_t>() : nullptr;
    int g = grid_for(n_rows * 32, 256, sm_count() * 16);
    constexpr int MAXQ = 4;  // 128 B-columns per pass
    for (int64_t b_off = 0; b_off < nB; b_off += 32 * MAXQ) {
        if (b_c_order)
            k_csr_dense<F, true, MAXQ><<<g, 256, 0, st>>>(data, indices, indptr, B, n, q, d, rows,
                                                          n_rows, ap, b_cols, nB, b_off, out);
        else
            k_csr_dense<F, false, MAXQ><<<g, 256, 0, st>>>(data, indices, indptr, B, n, q, d, rows,
                                                           n_rows, ap, b_cols, nB, b_off, out);
        TM_LAUNCHED();
    }
    return 0;
}

template <typename F>
int csr_matvec(const F* data, const int32_t* indices, const int32_t* indptr, int64_t n, int64_t p,
               const F* v, const int32_t* rows, int64_t n_rows, const int32_t* cols,
               int64_t n_cols, F* out, int accumulate, cudaStream_t st) {
    if (!rows) n_rows = n;
    if (n_rows <= 0) return 0;
    Scratch cmask((cols && n_cols < p) ? (size_t)p : 0, st);
    if (cmask.err != cudaSuccess) return fail_cuda(cmask.err, "scratch");
    const uint8_t* cm = nullptr;
    if (cols && n_cols < p) {
        int rc = build_mask(cols, n_cols, p, cmask.as<uint8_t>(), st);
        if (rc) return rc;
        cm = cmask.as<uint8_t>();
    }
    int g = grid_for(n_rows, 256, sm_count() * 32);
    k_csr_matvec<F><<<g, 256, 0, st>>>(data, indices, indptr, v, rows, n_rows, cm, out, accumulate);
    TM_LAUNCHED();
    return 0;
}

template <typename F, int MODE>
int csc_colreduce(const F* data, const int32_t* indices, const int32_t* indptr, int64_t n,
                  int64_t p, const F* v, const int32_t* rows, int64_t n_rows, const int32_t* cols,
                  int64_t n_cols, F* out, int accumulate, cudaStream_t st) {
    if (!cols) n_cols = p;
    if (n_cols <= 0) return 0;
    Scratch rmask((rows && n_rows < n) ? (size_t)n : 0, st);
    if (rmask.err != cudaSuccess) return fail_cuda(rmask.err, "scratch");
    const uint8_t* rm = nullptr;
    if (rows && n_rows < n) {
        int rc = build_mask(rows, n_rows, n, rmask.as<uint8_t>(), st);
        if (rc) return rc;
        rm = rmask.as<uint8_t>();
    }
    int g = grid_for(n_cols * 32, 256, sm_count() * 32);
    k_csc_colreduce<F, MODE><<<g, 256, 0, st>>>(data, indices, indptr, v, rm, cols, n_cols, out,
                                                accumulate);
    TM_LAUNCHED();
    return 0;
}

}  // namespace tmb

extern "C" {

#define TM_SPARSE_API(SUF, F)                                                                     \
    int tm_sparse_sandwich_##SUF(const F* csr_data, const int32_t* csr_indices,                   \
                                 const int32_t* csr_indptr, const int32_t* csr_row, int64_t n,    \
                                 int64_t p, int64_t nnz, const F* d, const int32_t* rows,         \
                                 int64_t n_rows, const int32_t* cols, int64_t n_cols, F* out,     \
                                 tm_stream_t stream) {                                            \
        return tmb::sparse_sandwich_ex<F>(csr_data, csr_indices, csr_indptr, csr_row, n, p, nnz,   \
                                          d, rows, n_rows, cols, n_cols, out,                     \
                                          tmb::as_stream(stream), false);                         \
    }                                                                                             \
    int tm_csr_dense_sandwich_##SUF(const F* csr_data, const int32_t* csr_indices,                \
                                    const int32_t* csr_indptr, int64_t n, int64_t p_sparse,       \
                                    const F* B, int64_t q, int b_c_order, const F* d,             \
                                    const int32_t* rows, int64_t n_rows, const int32_t* A_cols,   \
                                    int64_t nA, const int32_t* B_cols, int64_t nB, F* out,        \
                                    tm_stream_t stream) {                                         \
        return tmb::csr_dense_sandwich<F>(csr_data, csr_indices, csr_indptr, n, p_sparse, B, q,    \
                                         b_c_order, d, rows, n_rows, A_cols, nA, B_cols, nB, out, \
                                         tmb::as_stream(stream));                                  \
    }                                                                                             \
    int tm_csr_matvec_##SUF(const F* csr_data, const int32_t* csr_indices,                        \
                            const int32_t* csr_indptr, int64_t n, int64_t p, const F* v,          \
                            const int32_t* rows, int64_t n_rows, const int32_t* cols,             \
                            int64_t n_cols, F* out, int accumulate, tm_stream_t stream) {         \
        return tmb::csr_matvec<F>(csr_data, csr_indices, csr_indptr, n, p, v, rows, n_rows, cols,  \
                                 n_cols, out, accumulate, tmb::as_stream(stream));                 \
    }                                                                                             \
    int tm_csc_rmatvec_##SUF(const F* csc_data, const int32_t* csc_indices,                       \
                             const int32_t* csc_indptr, int64_t n, int64_t p, const F* v,         \
                             const int32_t* rows, int64_t n_rows, const int32_t* cols,            \
                             int64_t n_cols, F* out, int accumulate, tm_stream_t stream) {        \
        return tmb::csc_colreduce<F, 0>(csc_data, csc_indices, csc_indptr, n, p, v, rows, n_rows,  \
                                       cols, n_cols, out, accumulate, tmb::as_stream(stream));     \
    }                                                                                             \
    int tm_csc_sq_dot_weights_##SUF(const F* csc_data, const int32_t* csc_indices,                \
                                    const int32_t* csc_indptr, int64_t n, int64_t p, const F* w,  \
                                    F* out, tm_stream_t stream) {                                 \
        return tmb::csc_colreduce<F, 1>(csc_data, csc_indices, csc_indptr, n, p, w, nullptr, n,    \
                                       nullptr, p, out, 0, tmb::as_stream(stream));                \
    }

TM_SPARSE_API(f32, float)
TM_SPARSE_API(f64, double)

}  // extern "C"
