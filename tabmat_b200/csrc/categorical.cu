// Categorical block kernels: segmented reductions over the int32 category index.
// All of these are HBM-bound streams over (codes, weights) with a scatter into a small
// table; the table is privatised in shared memory when it fits.
//
// Reference semantics: categorical.pyx:23-218, split.pyx:32-111,
// cat_split_helpers-tmpl.cpp:4-151, categorical_matrix.py:825-838 (cat x sparse).
#include <cstdlib>

#include "tm_common.cuh"

namespace tmb {

constexpr int CAT_THREADS = 512;
constexpr int64_t CAT_SMEM_TABLE_BYTES = 40 * 1024;  // static-smem budget for a private table

// ---------------------------------------------------------------------------------------
// 1-D weighted histogram:  out[col(k)] += w[k]  for k in rows, col(k) = codes[k]-drop_first,
// optional column mask.  Serves cat sandwich (w=d) and cat transpose_matvec (w=v).
// ---------------------------------------------------------------------------------------
template <typename F, bool USE_SMEM>
__global__ void __launch_bounds__(CAT_THREADS)
k_cat_hist(const int32_t* __restrict__ codes, const F* __restrict__ w,
           const int32_t* __restrict__ rows, int64_t n_rows, int K, int drop_first,
           const uint8_t* __restrict__ col_mask, F* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* table = reinterpret_cast<F*>(smem_raw);
    if (USE_SMEM) {
        for (int i = threadIdx.x; i < K; i += blockDim.x) table[i] = F(0);
        __syncthreads();
    }
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += stride) {
        int64_t k = row_at(rows, t);
        int c = codes[k] - drop_first;
        if (c < 0) continue;
        if (col_mask && !col_mask[c]) continue;
        F val = w[k];
        if (USE_SMEM)
            atomicAdd(&table[c], val);
        else
            red_add(&out[c], val);
    }
    if (USE_SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            F v = table[i];
            if (v != F(0)) red_add(&out[i], v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Second-generation histogram kernels.  Differences to k_cat_hist / k_cat_cat above:
//  * a warp takes UNR groups of 32 CONSECUTIVE rows per visit and issues all their loads before
//    the first atomic (UNR x 32 x (4 + sizeof F) bytes in flight per warp);
//  * equal keys in consecutive lanes are summed in registers first (segmented suffix sum over
//    the maximal runs) and only the first lane of a run touches the table: on row-sorted
//    matrices (tabmat_b200/row_order.py) a whole warp collapses to one or two atomics, on
//    random codes the test costs one shuffle + one ballot;
//  * the shared-memory table is replicated as often as 48 KB allow (thread t adds into replica
//    t mod copies; a 10-level table ends up private to every thread): shared float atomics are
//    compare-and-swap loops on sm_100 (ATOMS.CAST.SPIN), so fewer lanes per address means
//    fewer retries.
// ---------------------------------------------------------------------------------------
template <typename F, typename KeyT>
__device__ __forceinline__ bool warp_run_reduce(KeyT key, F& val, int lane) {
    const unsigned FULL = 0xffffffffu;
    const KeyT prev = __shfl_up_sync(FULL, key, 1);
    const bool head = lane == 0 || prev != key;
    const unsigned heads = __ballot_sync(FULL, head);
    if (heads != FULL) {  // some run is longer than one lane
        const unsigned above = heads & ~((2u << lane) - 1u);  // run heads in lanes > lane
        const int end = above ? __ffs(above) - 2 : 31;        // last lane of this lane's run
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const F o = __shfl_down_sync(FULL, val, off);
            if (lane + off <= end) val += o;
        }
    }
    return head;
}

constexpr int CAT2_THREADS = 1024;

template <typename F, bool USE_SMEM, int UNR>
__global__ void __launch_bounds__(CAT2_THREADS, 1)
k_cat_hist2(const int32_t* __restrict__ codes, const F* __restrict__ w,
            const int32_t* __restrict__ rows, int64_t n_rows, int K, int drop_first,
            const uint8_t* __restrict__ col_mask, F* __restrict__ out, int copies) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* table = reinterpret_cast<F*>(smem_raw);
    if (USE_SMEM) {
        for (int i = threadIdx.x; i < K * copies; i += blockDim.x) table[i] = F(0);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    F* tab = USE_SMEM ? table + (threadIdx.x % copies) * K : out;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = gw * (32 * UNR); base < n_rows; base += nw * (32 * UNR)) {
        int c[UNR];
        F v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t t = base + u * 32 + lane;
            c[u] = -1;
            v[u] = F(0);
            if (t < n_rows) {
                const int64_t k = row_at(rows, t);
                c[u] = codes[k] - drop_first;
                v[u] = w[k];
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            int key = c[u];
            if (key >= 0 && col_mask && !col_mask[key]) key = -1;
            if (key < 0) {
                key = -1;
                v[u] = F(0);
            }
            const bool head = warp_run_reduce<F, int>(key, v[u], lane);
            if (head && key >= 0) {
                if (USE_SMEM)
                    atomicAdd(&tab[key], v[u]);
                else
                    red_add(&tab[key], v[u]);
            }
        }
    }
    if (USE_SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            F s = F(0);
            for (int r = 0; r < copies; ++r) s += table[r * K + i];
            if (s != F(0)) red_add(&out[i], s);
        }
    }
}

template <typename F, bool USE_SMEM, int UNR>
__global__ void __launch_bounds__(CAT2_THREADS, 1)
k_cat_cat2(const int32_t* __restrict__ ic, const int32_t* __restrict__ jc,
           const F* __restrict__ d, const int32_t* __restrict__ rows, int64_t n_rows, int Ki,
           int Kj, int dfi, int dfj, F* __restrict__ out, int copies) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* table = reinterpret_cast<F*>(smem_raw);
    const int tsz = Ki * Kj;
    if (USE_SMEM) {
        for (int i = threadIdx.x; i < tsz * copies; i += blockDim.x) table[i] = F(0);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    F* tab = USE_SMEM ? table + (threadIdx.x % copies) * tsz : out;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = gw * (32 * UNR); base < n_rows; base += nw * (32 * UNR)) {
        long long key[UNR];  // Ki * Kj may exceed 2^31 on the global-memory path
        F v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t t = base + u * 32 + lane;
            key[u] = -1;
            v[u] = F(0);
            if (t < n_rows) {
                const int64_t k = row_at(rows, t);
                const int i = ic[k] - dfi;
                const int j = jc[k] - dfj;
                if (i >= 0 && j >= 0) {
                    key[u] = (long long)i * Kj + j;
                    v[u] = d[k];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const bool head = warp_run_reduce<F, long long>(key[u], v[u], lane);
            if (head && key[u] >= 0) {
                if (USE_SMEM)
                    atomicAdd(&tab[key[u]], v[u]);
                else
                    red_add(&tab[key[u]], v[u]);
            }
        }
    }
    if (USE_SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < tsz; i += blockDim.x) {
            F s = F(0);
            for (int r = 0; r < copies; ++r) s += table[r * tsz + i];
            if (s != F(0)) red_add(&out[i], s);
        }
    }
}

// TABMAT_B200_CAT_V1=1 selects the first-generation kernels (k_cat_hist / k_cat_cat)
static bool cat_v1() {
    static const bool v1 = getenv("TABMAT_B200_CAT_V1") && atoi(getenv("TABMAT_B200_CAT_V1")) == 1;
    return v1;
}
constexpr int64_t CAT2_SMEM_BYTES = 48 * 1024;  // static limit: no opt-in attribute needed

static inline int cat2_copies(int64_t table_bytes) {
    int64_t c = CAT2_SMEM_BYTES / table_bytes;
    return (int)(c < 1 ? 1 : (c > CAT2_THREADS ? CAT2_THREADS : c));
}
// one fat CTA per SM; fewer when the input is small
static inline int cat2_grid(int64_t n_rows, int unr) {
    return grid_for(n_rows, CAT2_THREADS * unr, sm_count());
}

template <typename F>
int cat_hist2(const int32_t* codes, const F* w, const int32_t* rows, int64_t n_rows, int64_t K,
              int drop_first, const uint8_t* col_mask, F* out, bool overwrite, cudaStream_t st) {
    constexpr int UNR = 8;
    const int64_t tb = (int64_t)sizeof(F) * K;
    const int g = cat2_grid(n_rows, UNR);
    if (tb <= CAT_SMEM_TABLE_BYTES) {
        const int copies = cat2_copies(tb);
        if (overwrite) TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)K, st));
        k_cat_hist2<F, true, UNR><<<g, CAT2_THREADS, (size_t)(tb * copies), st>>>(
            codes, w, rows, n_rows, (int)K, drop_first, col_mask, out, copies);
    } else {
        if (overwrite) TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)K, st));
        k_cat_hist2<F, false, UNR><<<g, CAT2_THREADS, 0, st>>>(
            codes, w, rows, n_rows, (int)K, drop_first, col_mask, out, 1);
    }
    TM_LAUNCHED();
    return 0;
}

// `overwrite`: out = histogram (cat sandwich) instead of out += histogram (transpose_matvec)
template <typename F>
int cat_hist(const int32_t* codes, const F* w, const int32_t* rows, int64_t n_rows, int64_t K,
             int drop_first, const uint8_t* col_mask, F* out, cudaStream_t st,
             bool overwrite = false) {
    if (K <= 0) return 0;
    if (n_rows <= 0) {
        if (overwrite) TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)K, st));
        return 0;
    }
    if (!cat_v1())
        return cat_hist2<F>(codes, w, rows, n_rows, K, drop_first, col_mask, out, overwrite, st);
    if (overwrite) TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)K, st));
    bool smem = (int64_t)sizeof(F) * K <= CAT_SMEM_TABLE_BYTES;
    // few, fat CTAs when privatised (each flushes K bins); more CTAs otherwise
    int g = grid_for(n_rows, CAT_THREADS * 8, sm_count() * (smem ? 2 : 4));
    if (smem)
        k_cat_hist<F, true><<<g, CAT_THREADS, sizeof(F) * (size_t)K, st>>>(
            codes, w, rows, n_rows, (int)K, drop_first, col_mask, out);
    else
        k_cat_hist<F, false><<<g, CAT_THREADS, 0, st>>>(codes, w, rows, n_rows, (int)K, drop_first,
                                                        col_mask, out);
    TM_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// matvec (gather): out[i] += v[col(i)]
// ---------------------------------------------------------------------------------------
template <typename F>
__global__ void k_cat_matvec(const int32_t* __restrict__ codes, int64_t n,
                             const F* __restrict__ v, int drop_first,
                             const uint8_t* __restrict__ col_mask, F* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int c = codes[i] - drop_first;
        if (c < 0) continue;
        if (col_mask && !col_mask[c]) continue;
        out[i] += v[c];
    }
}

// ---------------------------------------------------------------------------------------
// cat x cat:  out[ci*Kj + cj] += d[k]
// ---------------------------------------------------------------------------------------
template <typename F, bool USE_SMEM>
__global__ void __launch_bounds__(CAT_THREADS)
k_cat_cat(const int32_t* __restrict__ ic, const int32_t* __restrict__ jc,
          const F* __restrict__ d, const int32_t* __restrict__ rows, int64_t n_rows, int Ki,
          int Kj, int dfi, int dfj, F* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* table = reinterpret_cast<F*>(smem_raw);
    const int tsz = Ki * Kj;
    if (USE_SMEM) {
        for (int i = threadIdx.x; i < tsz; i += blockDim.x) table[i] = F(0);
        __syncthreads();
    }
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += stride) {
        int64_t k = row_at(rows, t);
        int i = ic[k] - dfi;
        int j = jc[k] - dfj;
        if (i < 0 || j < 0) continue;
        F val = d[k];
        if (USE_SMEM)
            atomicAdd(&table[i * Kj + j], val);
        else
            red_add(&out[(int64_t)i * Kj + j], val);
    }
    if (USE_SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < tsz; i += blockDim.x) {
            F v = table[i];
            if (v != F(0)) red_add(&out[i], v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// cat x dense:  out[col(k)*nJ + b] += d[k] * Y[k, j_cols[b]]
// One warp per row; lanes stride over b, so every RED of a warp hits consecutive addresses.
// ---------------------------------------------------------------------------------------
template <typename F, bool C_ORDER>
__global__ void __launch_bounds__(256)
k_cat_dense(const int32_t* __restrict__ codes, const F* __restrict__ d,
            const F* __restrict__ Y, int64_t n, int64_t q, const int32_t* __restrict__ rows,
            int64_t n_rows, const int32_t* __restrict__ jcols, int64_t nJ, int drop_first,
            F* __restrict__ out) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp; t < n_rows; t += nwarps) {
        int64_t k = row_at(rows, t);
        int c = codes[k] - drop_first;
        if (c < 0) continue;
        F dk = d[k];
        F* orow = out + (int64_t)c * nJ;
        for (int64_t b = lane; b < nJ; b += 32) {
            int64_t j = jcols ? (int64_t)jcols[b] : b;
            F y = C_ORDER ? Y[k * q + j] : Y[j * n + k];
            red_add(&orow[b], dk * y);
        }
    }
}

// Privatised variant for small tables (K*nJ*sizeof(F) fits in shared memory):
// thread-owned columns, plain read-modify-write (no atomics) inside the CTA.
// blockDim.x = 128 threads own columns b = threadIdx.x (+128...), every thread walks all rows
// of the CTA's chunk.
template <typename F, bool C_ORDER>
__global__ void __launch_bounds__(128)
k_cat_dense_smem(const int32_t* __restrict__ codes, const F* __restrict__ d,
                 const F* __restrict__ Y, int64_t n, int64_t q,
                 const int32_t* __restrict__ rows, int64_t n_rows,
                 const int32_t* __restrict__ jcols, int nJ, int K, int drop_first,
                 F* __restrict__ out, int64_t rows_per_block) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* table = reinterpret_cast<F*>(smem_raw);  // [K][nJ]
    const int tsz = K * nJ;
    for (int i = threadIdx.x; i < tsz; i += blockDim.x) table[i] = F(0);
    __syncthreads();
    int64_t t0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t t1 = t0 + rows_per_block;
    if (t1 > n_rows) t1 = n_rows;
    for (int b = threadIdx.x; b < nJ; b += blockDim.x) {
        int64_t j = jcols ? (int64_t)jcols[b] : b;
        for (int64_t t = t0; t < t1; ++t) {
            int64_t k = row_at(rows, t);
            int c = codes[k] - drop_first;
            if (c < 0) continue;
            F y = C_ORDER ? Y[k * q + j] : Y[j * n + k];
            table[c * nJ + b] = fma(d[k], y, table[c * nJ + b]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tsz; i += blockDim.x) {
        F v = table[i];
        if (v != F(0)) red_add(&out[i], v);
    }
}

// ---------------------------------------------------------------------------------------
// cat x sparse (CSR + COO row):  out[col(k)*nS + pos(j)] += d[k] * A[k,j]
// one thread per non-zero
// ---------------------------------------------------------------------------------------
template <typename F>
__global__ void k_cat_sparse(const int32_t* __restrict__ codes, const F* __restrict__ d,
                             const F* __restrict__ data, const int32_t* __restrict__ indices,
                             const int32_t* __restrict__ nz_row, int64_t nnz,
                             const uint8_t* __restrict__ row_mask,
                             const int32_t* __restrict__ col_pos, int64_t nS, int drop_first,
                             F* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
        int k = nz_row[e];
        if (row_mask && !row_mask[k]) continue;
        int c = codes[k] - drop_first;
        if (c < 0) continue;
        int j = indices[e];
        int s = col_pos ? col_pos[j] : j;
        if (s < 0) continue;
        red_add(&out[(int64_t)c * nS + s], d[k] * data[e]);
    }
}

// ---- host wrappers ---------------------------------------------------------------------
template <typename F>
int cat_sandwich(const int32_t* codes, int64_t n, const F* d, const int32_t* rows, int64_t n_rows,
                 int64_t K, int drop_first, F* out, cudaStream_t st) {
    if (K <= 0) return 0;
    if (!rows) n_rows = n;
    return cat_hist<F>(codes, d, rows, n_rows, K, drop_first, nullptr, out, st, /*overwrite=*/true);
}

template <typename F>
int cat_transpose_matvec(const int32_t* codes, int64_t n, const F* v, const int32_t* rows,
                         int64_t n_rows, const int32_t* cols, int64_t n_cols, int64_t K,
                         int drop_first, F* out, cudaStream_t st) {
    if (K <= 0) return 0;
    if (!rows) n_rows = n;
    if (cols && n_cols < K) {
        Scratch mask((size_t)K, st);
        if (mask.err != cudaSuccess) return fail_cuda(mask.err, "scratch");
        int rc = build_mask(cols, n_cols, K, mask.as<uint8_t>(), st);
        if (rc) return rc;
        return cat_hist<F>(codes, v, rows, n_rows, K, drop_first, mask.as<uint8_t>(), out, st);
    }
    return cat_hist<F>(codes, v, rows, n_rows, K, drop_first, nullptr, out, st);
}

template <typename F>
int cat_matvec(const int32_t* codes, int64_t n, const F* v, const int32_t* cols, int64_t n_cols,
               int64_t K, int drop_first, F* out, cudaStream_t st) {
    if (n <= 0 || K <= 0) return 0;
    int g = grid_for(n, 256 * 4, sm_count() * 16);
    if (cols && n_cols < K) {
        Scratch mask((size_t)K, st);
        if (mask.err != cudaSuccess) return fail_cuda(mask.err, "scratch");
        int rc = build_mask(cols, n_cols, K, mask.as<uint8_t>(), st);
        if (rc) return rc;
        k_cat_matvec<F><<<g, 256, 0, st>>>(codes, n, v, drop_first, mask.as<uint8_t>(), out);
        TM_LAUNCHED();
        return 0;
    }
    k_cat_matvec<F><<<g, 256, 0, st>>>(codes, n, v, drop_first, nullptr, out);
    TM_LAUNCHED();
    return 0;
}

template <typename F>
int cat_cat_sandwich(const int32_t* ic, const int32_t* jc, int64_t n, int64_t Ki, int64_t Kj,
                     int dfi, int dfj, const F* d, const int32_t* rows, int64_t n_rows, F* out,
                     cudaStream_t st) {
    if (Ki <= 0 || Kj <= 0) return 0;
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)(Ki * Kj), st));
    if (!rows) n_rows = n;
    if (n_rows <= 0) return 0;
    bool smem = (int64_t)sizeof(F) * Ki * Kj <= CAT_SMEM_TABLE_BYTES;
    if (!cat_v1()) {
        constexpr int UNR = 4;
        const int64_t tb = (int64_t)sizeof(F) * Ki * Kj;
        const int g2 = cat2_grid(n_rows, UNR);
        if (smem) {
            const int copies = cat2_copies(tb);
            k_cat_cat2<F, true, UNR><<<g2, CAT2_THREADS, (size_t)(tb * copies), st>>>(
                ic, jc, d, rows, n_rows, (int)Ki, (int)Kj, dfi, dfj, out, copies);
        } else {
            k_cat_cat2<F, false, UNR><<<g2, CAT2_THREADS, 0, st>>>(
                ic, jc, d, rows, n_rows, (int)Ki, (int)Kj, dfi, dfj, out, 1);
        }
        TM_LAUNCHED();
        return 0;
    }
    int g = grid_for(n_rows, CAT_THREADS * 8, sm_count() * (smem ? 2 : 4));
    if (smem)
        k_cat_cat<F, true><<<g, CAT_THREADS, sizeof(F) * (size_t)(Ki * Kj), st>>>(
            ic, jc, d, rows, n_rows, (int)Ki, (int)Kj, dfi, dfj, out);
    else
        k_cat_cat<F, false><<<g, CAT_THREADS, 0, st>>>(ic, jc, d, rows, n_rows, (int)Ki, (int)Kj,
                                                       dfi, dfj, out);
    TM_LAUNCHED();
    return 0;
}

template <typename F>
int cat_dense_sandwich(const int32_t* codes, int64_t n, int64_t K, int drop_first, const F* d,
                       const F* Y, int64_t q, int y_c_order, const int32_t* rows, int64_t n_rows,
                       const int32_t* jcols, int64_t nJ, F* out, cudaStream_t st) {
    if (!jcols) nJ = q;
    if (K <= 0 || nJ <= 0) return 0;
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)(K * nJ), st));
    if (!rows) n_rows = n;
    if (n_rows <= 0) return 0;
    int64_t table_bytes = (int64_t)sizeof(F) * K * nJ;
    if (table_bytes <= CAT_SMEM_TABLE_BYTES && nJ >= 64) {
        int64_t blocks = (int64_t)sm_count() * 4;
        int64_t rpb = (n_rows + blocks - 1) / blocks;
        if (rpb < 64) rpb = 64;
        blocks = (n_rows + rpb - 1) / rpb;
        if (y_c_order)
            k_cat_dense_smem<F, true><<<(unsigned)blocks, 128, (size_t)table_bytes, st>>>(
                codes, d, Y, n, q, rows, n_rows, jcols, (int)nJ, (int)K, drop_first, out, rpb);
        else
            k_cat_dense_smem<F, false><<<(unsigned)blocks, 128, (size_t)table_bytes, st>>>(
                codes, d, Y, n, q, rows, n_rows, jcols, (int)nJ, (int)K, drop_first, out, rpb);
    } else {
        int g = grid_for(n_rows * 32, 256, sm_count() * 16);
        if (y_c_order)
            k_cat_dense<F, true><<<g, 256, 0, st>>>(codes, d, Y, n, q, rows, n_rows, jcols, nJ,
                                                    drop_first, out);
        else
            k_cat_dense<F, false><<<g, 256, 0, st>>>(codes, d, Y, n, q, rows, n_rows, jcols, nJ,
                                                     drop_first, out);
    }
    TM_LAUNCHED();
    return 0;
}

template <typename F>
int cat_sparse_sandwich(const int32_t* codes, int64_t n, int64_t K, int drop_first, const F* d,
                        const F* data, const int32_t* indices, const int32_t* indptr,
                        const int32_t* nz_row, int64_t p_sparse, int64_t nnz, const int32_t* rows,
                        int64_t n_rows, const int32_t* s_cols, int64_t nS, F* out,
                        cudaStream_t st) {
    (void)indptr;
    if (!s_cols) nS = p_sparse;
    if (K <= 0 || nS <= 0) return 0;
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)(K * nS), st));
    if (nnz <= 0) return 0;
    if (rows && n_rows <= 0) return 0;
    Scratch rmask(rows ? (size_t)n : 0, st);
    Scratch cpos(s_cols ? sizeof(int32_t) * (size_t)p_sparse : 0, st);
    if (rmask.err != cudaSuccess) return fail_cuda(rmask.err, "scratch");
    if (cpos.err != cudaSuccess) return fail_cuda(cpos.err, "scratch");
    if (rows) {
        int rc = build_mask(rows, n_rows, n, rmask.as<uint8_t>(), st);
        if (rc) return rc;
    }
    if (s_cols) {
        int rc = build_pos_map(s_cols, nS, p_sparse, cpos.as<int32_t>(), st);
        if (rc) return rc;
    }
    int g = grid_for(nnz, 256 * 4, sm_count() * 16);
    k_cat_sparse<F><<<g, 256, 0, st>>>(codes, d, data, indices, nz_row, nnz,
                                       rows ? rmask.as<uint8_t>() : nullptr,
                                       s_cols ? cpos.as<int32_t>() : nullptr, nS, drop_first, out);
    TM_LAUNCHED();
    return 0;
}

}  // namespace tmb

extern "C" {

#define TM_CAT_API(SUF, F)                                                                        \
    int tm_cat_sandwich_##SUF(const int32_t* codes, int64_t n, const F* d, const int32_t* rows,   \
                              int64_t n_rows, int64_t K, int drop_first, F* out,                  \
                              tm_stream_t stream) {                                               \
        return tmb::cat_sandwich<F>(codes, n, d, rows, n_rows, K, drop_first, out,                 \
                                   tmb::as_stream(stream));                                        \
    }                                                                                             \
    int tm_cat_transpose_matvec_##SUF(const int32_t* codes, int64_t n, const F* v,                \
                                      const int32_t* rows, int64_t n_rows, const int32_t* cols,   \
                                      int64_t n_cols, int64_t K, int drop_first, F* out,          \
                                      tm_stream_t stream) {                                       \
        return tmb::cat_transpose_matvec<F>(codes, n, v, rows, n_rows, cols, n_cols, K,            \
                                           drop_first, out, tmb::as_stream(stream));               \
    }                                                                                             \
    int tm_cat_matvec_##SUF(const int32_t* codes, int64_t n, const F* v, const int32_t* cols,     \
                            int64_t n_cols, int64_t K, int drop_first, F* out,                    \
                            tm_stream_t stream) {                                                 \
        return tmb::cat_matvec<F>(codes, n, v, cols, n_cols, K, drop_first, out,                   \
                                 tmb::as_stream(stream));                                          \
    }                                                                                             \
    int tm_cat_dense_sandwich_##SUF(const int32_t* codes, int64_t n, int64_t K, int drop_first,   \
                                    const F* d, const F* Y, int64_t q, int y_c_order,             \
                                    const int32_t* rows, int64_t n_rows, const int32_t* j_cols,   \
                                    int64_t nJ, F* out, tm_stream_t stream) {                     \
        return tmb::cat_dense_sandwich<F>(codes, n, K, drop_first, d, Y, q, y_c_order, rows,       \
                                         n_rows, j_cols, nJ, out, tmb::as_stream(stream));         \
    }                                                                                             \
    int tm_cat_cat_sandwich_##SUF(const int32_t* ic, const int32_t* jc, int64_t n, int64_t Ki,    \
                                  int64_t Kj, int dfi, int dfj, const F* d, const int32_t* rows,  \
                                  int64_t n_rows, F* out, tm_stream_t stream) {                   \
        return tmb::cat_cat_sandwich<F>(ic, jc, n, Ki, Kj, dfi, dfj, d, rows, n_rows, out,         \
                                       tmb::as_stream(stream));                                    \
    }                                                                                             \
    int tm_cat_sparse_sandwich_##SUF(                                                             \
        const int32_t* codes, int64_t n, int64_t K, int drop_first, const F* d,                   \
        const F* csr_data, const int32_t* csr_indices, const int32_t* csr_indptr,                 \
        const int32_t* csr_row, int64_t p_sparse, int64_t nnz, const int32_t* rows,               \
        int64_t n_rows, const int32_t* s_cols, int64_t nS, F* out, tm_stream_t stream) {          \
        return tmb::cat_sparse_sandwich<F>(codes, n, K, drop_first, d, csr_data, csr_indices,      \
                                          csr_indptr, csr_row, p_sparse, nnz, rows, n_rows,       \
                                          s_cols, nS, out, tmb::as_stream(stream));                \
    }

TM_CAT_API(f32, float)
TM_CAT_API(f64, double)

}  // extern "C"
