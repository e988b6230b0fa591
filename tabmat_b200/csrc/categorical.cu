// Categorical block kernels: segmented reductions over the int32 category index.
// All of these are HBM-bound streams over (codes, weights) with a scatter into a small
// table; the table is privatised in shared memory when it fits.
//
// Reference semantics: categorical.pyx:23-218, split.pyx:32-111,
// cat_split_helpers-tmpl.cpp:4-151, categorical_matrix.py:825-838 (cat x sparse).
#include <cstdlib>

#include <cub/device/device_scan.cuh>

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

// ---------------------------------------------------------------------------------------
// Third generation, the unrestricted case (all rows, no column mask, table in shared memory):
// the shape tools/micro/hist_bench.cu found fastest for an 80 MB one-shot stream (25.5 us
// against 25.6 us for a kernel that only reads the 80 MB): 16-byte loads (4 consecutive rows per
// thread), V of them in flight per thread, two CTAs of 1024 threads per SM.  Equal codes in
// consecutive rows are merged inside the thread first, and a warp whose 128 rows all carry the
// same code (row-sorted storage) adds them with one shuffle tree and ONE atomic.
// ---------------------------------------------------------------------------------------
template <typename F>
struct Ld4;
template <>
struct Ld4<float> {
    static __device__ __forceinline__ void load(const float* w, int64_t t4, float (&v)[4]) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(w) + t4);
        v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w;
    }
};
template <>
struct Ld4<double> {
    static __device__ __forceinline__ void load(const double* w, int64_t t4, double (&v)[4]) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(w) + 2 * t4);
        const double2 b = __ldg(reinterpret_cast<const double2*>(w) + 2 * t4 + 1);
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
    }
};

template <typename F, int V>
__global__ void __launch_bounds__(CAT2_THREADS, 2)
k_cat_hist_vec(const int32_t* __restrict__ codes, const F* __restrict__ w, int64_t n, int K,
               int drop_first, F* __restrict__ out, int copies) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* table = reinterpret_cast<F*>(smem_raw);
    for (int i = threadIdx.x; i < K * copies; i += CAT2_THREADS) table[i] = F(0);
    __syncthreads();
    F* tab = table + (threadIdx.x % copies) * K;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * CAT2_THREADS;
    // all threads of a warp run the same number of iterations (t differs by < 32)
    const int64_t t_warp0 = (int64_t)blockIdx.x * CAT2_THREADS + (threadIdx.x & ~31);
    for (int64_t tw = t_warp0; tw < n4; tw += stride * V) {
        int4 c[V];
        F v[V][4];
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int64_t tt = tw + lane + u * stride;
            c[u] = make_int4(-1, -1, -1, -1);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[u][i] = F(0);
            if (tt < n4) {
                c[u] = __ldg(reinterpret_cast<const int4*>(codes) + tt);
                Ld4<F>::load(w, tt, v[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
            int k0 = c[u].x - drop_first, k1 = c[u].y - drop_first, k2 = c[u].z - drop_first,
                k3 = c[u].w - drop_first;
            if (c[u].x < 0) k0 = -1;   // rows past the end
            // merge equal neighbours: the sum travels to the last row of the run
            if (k1 == k0) { v[u][1] += v[u][0]; k0 = -1; }
            if (k2 == k1) { v[u][2] += v[u][1]; k1 = -1; }
            if (k3 == k2) { v[u][3] += v[u][2]; k2 = -1; }
            // the whole warp on one code (row-sorted storage): one atomic for 128 rows
            const bool one = k0 < 0 && k1 < 0 && k2 < 0;
            const int kf = __shfl_sync(FULL, k3, 0);
            if (__all_sync(FULL, one && k3 == kf)) {
                F sum = v[u][3];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
                if (lane == 0 && kf >= 0) atomicAdd(&tab[kf], sum);
                continue;
            }
            if (k0 >= 0) atomicAdd(&tab[k0], v[u][0]);
            if (k1 >= 0) atomicAdd(&tab[k1], v[u][1]);
            if (k2 >= 0) atomicAdd(&tab[k2], v[u][2]);
            if (k3 >= 0) atomicAdd(&tab[k3], v[u][3]);
        }
    }
    // the n % 4 last rows
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t k = (n4 << 2) + threadIdx.x;
        const int key = codes[k] - drop_first;
        if (key >= 0) atomicAdd(&tab[key], w[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += CAT2_THREADS) {
        F s = F(0);
        for (int r = 0; r < copies; ++r) s += table[r * K + i];
        if (s != F(0)) red_add(&out[i], s);
    }
}

static bool cat_vec_off() {
    static const bool off = getenv("TABMAT_B200_CAT_VEC") && atoi(getenv("TABMAT_B200_CAT_VEC")) == 0;
    return off;
}

template <typename F>
int cat_hist2(const int32_t* codes, const F* w, const int32_t* rows, int64_t n_rows, int64_t K,
              int drop_first, const uint8_t* col_mask, F* out, bool overwrite, cudaStream_t st) {
    constexpr int64_t VEC_TABLE_BYTES = 32 * 1024;   // x copies <= 96 KB per CTA, two CTAs per SM
    const int64_t tbv = (int64_t)sizeof(F) * K;
    if (!rows && !col_mask && !cat_vec_off() && tbv <= VEC_TABLE_BYTES && n_rows >= 4096 &&
        ((reinterpret_cast<uintptr_t>(codes) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
        constexpr int V = sizeof(F) == 4 ? 2 : 1;   // 32 registers per thread at 2 x 1024 threads
        int64_t cp = (96 * 1024) / tbv;
        const int copies = (int)(cp < 1 ? 1 : (cp > 32 ? 32 : cp));
        const size_t smem = (size_t)(tbv * copies);
        TM_CUDA(cudaFuncSetAttribute(k_cat_hist_vec<F, V>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        if (overwrite) TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)K, st));
        const int g = grid_for(n_rows >> 2, CAT2_THREADS * V, sm_count() * 2);
        k_cat_hist_vec<F, V><<<g, CAT2_THREADS, smem, st>>>(codes, w, n_rows, (int)K, drop_first,
                                                            out, copies);
        TM_LAUNCHED();
        return 0;
    }
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

// unrestricted case with 16-byte accesses: 4 consecutive rows per thread, two loads in flight
template <typename F>
__global__ void __launch_bounds__(256)
k_cat_matvec_vec(const int32_t* __restrict__ codes, int64_t n, const F* __restrict__ v,
                 int drop_first, F* __restrict__ out) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += 2 * stride) {
        const int64_t t2 = t + stride;
        const bool two = t2 < n4;
        int4 c[2];
        F o[2][4];
        c[0] = __ldg(reinterpret_cast<const int4*>(codes) + t);
        if (two) c[1] = __ldg(reinterpret_cast<const int4*>(codes) + t2);
        Ld4<F>::load(out, t, o[0]);
        if (two) Ld4<F>::load(out, t2, o[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const int k0 = c[u].x - drop_first, k1 = c[u].y - drop_first, k2 = c[u].z - drop_first,
                      k3 = c[u].w - drop_first;
            if (k0 >= 0) o[u][0] += __ldg(v + k0);
            if (k1 >= 0) o[u][1] += __ldg(v + k1);
            if (k2 >= 0) o[u][2] += __ldg(v + k2);
            if (k3 >= 0) o[u][3] += __ldg(v + k3);
            F* dst = out + 4 * (u == 0 ? t : t2);
            if (sizeof(F) == 4) {
                *reinterpret_cast<float4*>(dst) = make_float4((float)o[u][0], (float)o[u][1],
                                                              (float)o[u][2], (float)o[u][3]);
            } else {
                reinterpret_cast<double2*>(dst)[0] = make_double2((double)o[u][0], (double)o[u][1]);
                reinterpret_cast<double2*>(dst)[1] = make_double2((double)o[u][2], (double)o[u][3]);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        const int c = codes[i] - drop_first;
        if (c >= 0) out[i] += v[c];
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

// ---------------------------------------------------------------------------------------
// CSR form of a categorical block on the device: multiply_complex / subset_categorical_complex
// (categorical.pyx:221-315) without the host round trip.  flags -> exclusive scan (CUB) ->
// compaction; row i owns at most one entry (column codes[i] - drop_first, value d[i] or 1).
// ---------------------------------------------------------------------------------------
__global__ void k_cat_flags(const int32_t* __restrict__ codes, int64_t n, int drop_first,
                            int32_t* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride)
        flags[i] = (i < n && codes[i] >= drop_first) ? 1 : 0;
}
template <typename F>
__global__ void k_cat_compact(const int32_t* __restrict__ codes, int64_t n, int drop_first,
                              const F* __restrict__ d, const int32_t* __restrict__ indptr,
                              F* __restrict__ data, int32_t* __restrict__ indices) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = codes[i] - drop_first;
        if (c < 0) continue;
        const int32_t pos = indptr[i];
        indices[pos] = c;
        if (data) data[pos] = d ? d[i] : F(1);
    }
}

template <typename F>
int cat_to_csr(const int32_t* codes, int64_t n, int drop_first, const F* d, F* data,
               int32_t* indices, int32_t* indptr, cudaStream_t st) {
    if (n < 0 || n >= 0x7fffffffLL) return fail("tm_cat_to_csr: n out of range");
    Scratch flags(sizeof(int32_t) * (size_t)(n + 1), st);
    if (flags.err != cudaSuccess) return fail_cuda(flags.err, "scratch");
    const int g = grid_for(n + 1, 256 * 4, sm_count() * 16);
    k_cat_flags<<<g, 256, 0, st>>>(codes, n, drop_first, flags.as<int32_t>());
    TM_LAUNCHED();
    size_t tmp_bytes = 0;
    TM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags.as<int32_t>(), indptr,
                                          (int)(n + 1), st));
    Scratch tmp(tmp_bytes, st);
    if (tmp.err != cudaSuccess) return fail_cuda(tmp.err, "scratch");
    TM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, flags.as<int32_t>(), indptr,
                                          (int)(n + 1), st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (n > 0) {
        k_cat_compact<F><<<g, 256, 0, st>>>(codes, n, drop_first, d, indptr, data, indices);
        TM_LAUNCHED();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// Deterministic weighted histogram (opt-in, tm_set_deterministic): out[c] (+)= sum of w[k] over
// the rows of category c, added in a FIXED order — the rows of a category come from the cached
// sorted permutation (perm / segptr, like the sorted-gather kernel), one warp per category, lane
// l adds entries l, l+32, ... in sequence, then a fixed butterfly.  Bit-identical from run to
// run, unlike the atomic kernels above (the reference made categorical transpose_matvec
// deterministic on purpose, CHANGELOG.rst:134).  Gathers 4-8 bytes per row: slower, opt-in.
// `row_w` (nullable) = 0/1 row mask folded in as a factor.
// ---------------------------------------------------------------------------------------
template <typename F>
__global__ void __launch_bounds__(256)
k_cat_segment_sum(const F* __restrict__ w, const F* __restrict__ row_w,
                  const int32_t* __restrict__ perm, const int32_t* __restrict__ segptr, int K,
                  F* __restrict__ out, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    for (int c = warp; c < K; c += nwarps) {
        const int e0 = segptr[c], e1 = segptr[c + 1];
        F s = F(0);
        for (int e = e0 + lane; e < e1; e += 32) {
            const int k = perm[e];
            s += row_w ? w[k] * row_w[k] : w[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[c] = accumulate ? out[c] + s : s;
    }
}

template <typename F>
int cat_segment_sum(const F* w, const F* row_w, const int32_t* perm, const int32_t* segptr,
                    int64_t K, F* out, int accumulate, cudaStream_t st) {
    if (K <= 0) return 0;
    const int g = grid_for(K * 32, 256, sm_count() * 8);
    k_cat_segment_sum<F><<<g, 256, 0, st>>>(w, row_w, perm, segptr, (int)K, out, accumulate);
    TM_LAUNCHED();
    return 0;
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
    if (!cat_vec_off() && n >= 4096 &&
        ((reinterpret_cast<uintptr_t>(codes) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const int gv = grid_for(n >> 2, 256 * 2, sm_count() * 16);
        k_cat_matvec_vec<F><<<gv, 256, 0, st>>>(codes, n, v, drop_first, out);
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
    int tm_cat_segment_sum_##SUF(const F* w, const F* row_w, const int32_t* perm,                 \
                                 const int32_t* segptr, int64_t K, F* out, int accumulate,        \
                                 tm_stream_t stream) {                                            \
        return tmb::cat_segment_sum<F>(w, row_w, perm, segptr, K, out, accumulate,                 \
                                      tmb::as_stream(stream));                                     \
    }                                                                                             \
    int tm_cat_to_csr_##SUF(const int32_t* codes, int64_t n, int drop_first, const F* d, F* data, \
                            int32_t* indices, int32_t* indptr, tm_stream_t stream) {              \
        return tmb::cat_to_csr<F>(codes, n, drop_first, d, data, indices, indptr,                  \
                                 tmb::as_stream(stream));                                          \
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
