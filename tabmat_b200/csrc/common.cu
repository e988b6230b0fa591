// Library-level state, small utility kernels, SplitMatrix block assembly.
#include "tm_common.cuh"

namespace tmb {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void keep_pool_memory() {
    static thread_local int done_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_dev = dev;
}

int sm_count() {
    // cached per device: a process may drive several GPUs
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n) return n;
    n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
    return n;
}

__global__ void k_build_pos_map(const int32_t* __restrict__ list, int64_t m,
                                int32_t* __restrict__ map) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) map[list[i]] = (int32_t)i;
}

__global__ void k_build_mask(const int32_t* __restrict__ list, int64_t m,
                             uint8_t* __restrict__ mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) mask[list[i]] = 1;
}

int build_pos_map(const int32_t* list, int64_t m, int64_t p, int32_t* map, cudaStream_t st) {
    TM_CUDA(cudaMemsetAsync(map, 0xFF, sizeof(int32_t) * (size_t)p, st));  // -1
    if (m > 0) {
        k_build_pos_map<<<grid_for(m, 256, kMaxGridX), 256, 0, st>>>(list, m, map);
        TM_LAUNCHED();
    }
    return 0;
}

int build_mask(const int32_t* list, int64_t m, int64_t p, uint8_t* mask, cudaStream_t st) {
    TM_CUDA(cudaMemsetAsync(mask, 0, (size_t)p, st));
    if (m > 0) {
        k_build_mask<<<grid_for(m, 256, kMaxGridX), 256, 0, st>>>(list, m, mask);
        TM_LAUNCHED();
    }
    return 0;
}

template <typename F>
__global__ void k_masked_weights(const F* __restrict__ d, const int32_t* __restrict__ rows,
                                 int64_t n_rows, F* __restrict__ dmask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_rows; i += stride) {
        int32_t k = rows[i];
        // a row listed twice counts twice, like the index-list gather of the passes that walk
        // `rows` directly (the reference's dense kernels treat `rows` as a multiset)
        red_add(dmask + k, d[k]);
    }
}

template <typename F>
int masked_weights(const F* d, int64_t n, const int32_t* rows, int64_t n_rows, F* dmask,
                   cudaStream_t st) {
    TM_CUDA(cudaMemsetAsync(dmask, 0, sizeof(F) * (size_t)n, st));
    if (n_rows > 0) {
        k_masked_weights<F><<<grid_for(n_rows, 256, sm_count() * 16), 256, 0, st>>>(d, rows, n_rows,
                                                                                    dmask);
        TM_LAUNCHED();
    }
    return 0;
}
template int masked_weights<float>(const float*, int64_t, const int32_t*, int64_t, float*,
                                   cudaStream_t);
template int masked_weights<double>(const double*, int64_t, const int32_t*, int64_t, double*,
                                    cudaStream_t);

template <typename F, bool UPPER_SRC>
__global__ void k_symmetrize(F* __restrict__ out, int64_t m) {
    // 32x32 tiles; source tiles are on/below the diagonal (or on/above when UPPER_SRC)
    __shared__ F tile[32][33];
    int64_t tr = blockIdx.y, tc = blockIdx.x;
    if (UPPER_SRC ? (tc < tr) : (tc > tr)) return;
    int64_t r = tr * 32 + threadIdx.y, c = tc * 32 + threadIdx.x;
    for (int i = 0; i < 32; i += 8) {
        int64_t rr = r + i;
        if (rr < m && c < m) tile[threadIdx.y + i][threadIdx.x] = out[rr * m + c];
    }
    __syncthreads();
    // write transposed: element (c', r') = tile[r'][c']
    int64_t r2 = tc * 32 + threadIdx.y, c2 = tr * 32 + threadIdx.x;
    for (int i = 0; i < 32; i += 8) {
        int64_t rr = r2 + i;  // destination row (source column)
        bool in_dst = UPPER_SRC ? (c2 < rr) : (c2 > rr);
        if (rr < m && c2 < m && in_dst) out[rr * m + c2] = tile[threadIdx.x][threadIdx.y + i];
    }
}

template <typename F>
int symmetrize_from_lower(F* out, int64_t m, cudaStream_t st) {
    if (m <= 1) return 0;
    int t = (int)((m + 31) / 32);
    dim3 grid(t, t), block(32, 8);
    k_symmetrize<F, false><<<grid, block, 0, st>>>(out, m);
    TM_LAUNCHED();
    return 0;
}
template int symmetrize_from_lower<float>(float*, int64_t, cudaStream_t);
template int symmetrize_from_lower<double>(double*, int64_t, cudaStream_t);

template <typename F>
int symmetrize_from_upper(F* out, int64_t m, cudaStream_t st) {
    if (m <= 1) return 0;
    int t = (int)((m + 31) / 32);
    dim3 grid(t, t), block(32, 8);
    k_symmetrize<F, true><<<grid, block, 0, st>>>(out, m);
    TM_LAUNCHED();
    return 0;
}
template int symmetrize_from_upper<float>(float*, int64_t, cudaStream_t);
template int symmetrize_from_upper<double>(double*, int64_t, cudaStream_t);

// ---- SplitMatrix block placement (split_matrix.py:336-354) -----------------------------
template <typename F>
__global__ void k_scatter_block(const F* __restrict__ blk, int64_t na, int64_t nb,
                                const int64_t* __restrict__ ri, const int64_t* __restrict__ ci,
                                double* __restrict__ out, int64_t ld, int mirror) {
    __shared__ F tile[32][33];
    int64_t a0 = (int64_t)blockIdx.y * 32, b0 = (int64_t)blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int64_t a = a0 + i, b = b0 + threadIdx.x;
        if (a < na && b < nb) {
            F v = blk[a * nb + b];
            tile[i][threadIdx.x] = v;
            int64_t r = ri ? ri[a] : a, c = ci ? ci[b] : b;
            if (r >= 0 && c >= 0) out[r * ld + c] = (double)v;   // negative = column not selected
        }
    }
    if (!mirror) return;
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int64_t b = b0 + i, a = a0 + threadIdx.x;
        if (a < na && b < nb) {
            int64_t r = ri ? ri[a] : a, c = ci ? ci[b] : b;
            if (r >= 0 && c >= 0) out[c * ld + r] = (double)tile[threadIdx.x][i];
        }
    }
}

template <typename F>
__global__ void k_scatter_diag(const F* __restrict__ diag, int64_t na,
                               const int64_t* __restrict__ ri, double* __restrict__ out,
                               int64_t ld) {
    int64_t a = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (a >= na) return;
    int64_t r = ri ? ri[a] : a;
    if (r < 0) return;
    for (int64_t b = (int64_t)blockIdx.x * 32 + threadIdx.x; b < na; b += (int64_t)gridDim.x * 32) {
        int64_t c = ri ? ri[b] : b;
        if (c >= 0) out[r * ld + c] = (a == b) ? (double)diag[a] : 0.0;
    }
}

template <typename F>
int scatter_block(const F* blk, int64_t na, int64_t nb, const int64_t* ri, const int64_t* ci,
                  double* out, int64_t ld, int mirror, cudaStream_t st) {
    if (na <= 0 || nb <= 0) return 0;
    dim3 grid((unsigned)((nb + 31) / 32), (unsigned)((na + 31) / 32)), block(32, 8);
    k_scatter_block<F><<<grid, block, 0, st>>>(blk, na, nb, ri, ci, out, ld, mirror);
    TM_LAUNCHED();
    return 0;
}

template <typename F>
int scatter_diag(const F* diag, int64_t na, const int64_t* ri, double* out, int64_t ld,
                 cudaStream_t st) {
    if (na <= 0) return 0;
    int gx = (int)((na + 31) / 32);
    if (gx > 64) gx = 64;
    dim3 grid(gx, (unsigned)((na + 7) / 8)), block(32, 8);
    k_scatter_diag<F><<<grid, block, 0, st>>>(diag, na, ri, out, ld);
    TM_LAUNCHED();
    return 0;
}

// ---- row-order permutation of length-n vectors (tabmat_b200/row_order.py) ---------------
// GATHER: dst[i] = src[perm[i]]   SCATTER: dst[perm[i]] = src[i] (or += when accumulate)
template <typename F, bool GATHER>
__global__ void k_permute(const F* __restrict__ src, const int32_t* __restrict__ perm, int64_t n,
                          F* __restrict__ dst, int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t j = perm[i];
        if (GATHER)
            dst[i] = accumulate ? dst[i] + src[j] : src[j];
        else
            dst[j] = accumulate ? dst[j] + src[i] : src[i];
    }
}

template <typename F>
int permute(const F* src, const int32_t* perm, int64_t n, F* dst, int gather, int accumulate,
            cudaStream_t st) {
    if (n <= 0) return 0;
    int g = grid_for(n, 256 * 4, sm_count() * 16);
    if (gather)
        k_permute<F, true><<<g, 256, 0, st>>>(src, perm, n, dst, accumulate);
    else
        k_permute<F, false><<<g, 256, 0, st>>>(src, perm, n, dst, accumulate);
    TM_LAUNCHED();
    return 0;
}

// ---- StandardizedMatrix.sandwich epilogue (standardized_mat.py:123-172) in one kernel ------
//   out[i, j] = term1[i, j] * mult[i] * mult[j] + dm[i] * shift[j] + shift[i] * dm[j]
//               + shift[i] * shift[j] * sum_d,        dm[i] = d_mat[i] * mult[i]
// term1 = the inner matrix' sandwich (float64 for a SplitMatrix, else F; a categorical's diagonal
// when `diag`), d_mat = inner.transpose_matvec(d), sum_d a DEVICE scalar (no host sync).
template <typename TI, typename F>
__global__ void k_std_combine(const TI* __restrict__ term1, int diag, const F* __restrict__ dmat,
                              const F* __restrict__ shift, const F* __restrict__ mult,
                              const F* __restrict__ sum_d, int64_t m, F* __restrict__ out) {
    const int64_t total = m * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const F sd = *sum_d;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t i = t / m, j = t - i * m;
        const F mi = mult ? mult[i] : F(1), mj = mult ? mult[j] : F(1);
        const F si = shift[i], sj = shift[j];
        F r = dmat[i] * mi * sj + si * (dmat[j] * mj) + si * sj * sd;
        if (diag) {
            if (i == j) r += (F)term1[i] * mi * mj;
        } else {
            r += (F)term1[t] * (mi * mj);
        }
        out[t] = r;
    }
}

template <typename F>
int std_combine(const void* term1, int term1_f64, int diag, const F* dmat, const F* shift,
                const F* mult, const F* sum_d, int64_t m, F* out, cudaStream_t st) {
    if (m <= 0) return 0;
    const int g = grid_for(m * m, 256 * 2, sm_count() * 8);
    if (term1_f64)
        k_std_combine<double, F><<<g, 256, 0, st>>>(static_cast<const double*>(term1), diag, dmat,
                                                   shift, mult, sum_d, m, out);
    else
        k_std_combine<F, F><<<g, 256, 0, st>>>(static_cast<const F*>(term1), diag, dmat, shift,
                                              mult, sum_d, m, out);
    TM_LAUNCHED();
    return 0;
}

}  // namespace tmb

extern "C" {

int tm_std_sandwich_combine_f32(const void* term1, int term1_f64, int diag, const float* dmat,
                                const float* shift, const float* mult, const float* sum_d,
                                int64_t m, float* out, tm_stream_t stream) {
    return tmb::std_combine<float>(term1, term1_f64, diag, dmat, shift, mult, sum_d, m, out,
                                   tmb::as_stream(stream));
}
int tm_std_sandwich_combine_f64(const void* term1, int term1_f64, int diag, const double* dmat,
                                const double* shift, const double* mult, const double* sum_d,
                                int64_t m, double* out, tm_stream_t stream) {
    return tmb::std_combine<double>(term1, 1, diag, dmat, shift, mult, sum_d, m, out,
                                    tmb::as_stream(stream));
}

int tm_permute_gather_f32(const float* src, const int32_t* perm, int64_t n, float* dst,
                          int accumulate, tm_stream_t stream) {
    return tmb::permute<float>(src, perm, n, dst, 1, accumulate, tmb::as_stream(stream));
}
int tm_permute_gather_f64(const double* src, const int32_t* perm, int64_t n, double* dst,
                          int accumulate, tm_stream_t stream) {
    return tmb::permute<double>(src, perm, n, dst, 1, accumulate, tmb::as_stream(stream));
}
int tm_permute_scatter_f32(const float* src, const int32_t* perm, int64_t n, float* dst,
                           int accumulate, tm_stream_t stream) {
    return tmb::permute<float>(src, perm, n, dst, 0, accumulate, tmb::as_stream(stream));
}
int tm_permute_scatter_f64(const double* src, const int32_t* perm, int64_t n, double* dst,
                           int accumulate, tm_stream_t stream) {
    return tmb::permute<double>(src, perm, n, dst, 0, accumulate, tmb::as_stream(stream));
}

int tm_version(void) { return 100; }
const char* tm_last_error(void) { return tmb::g_err; }
int64_t tm_launch_count(void) { return (int64_t)tmb::g_launches.load(); }
void tm_reset_launch_count(void) { tmb::g_launches.store(0); }

int tm_scatter_block_f32(const float* blk, int64_t na, int64_t nb, const int64_t* ri,
                         const int64_t* ci, double* out, int64_t ld, int mirror,
                         tm_stream_t stream) {
    return tmb::scatter_block<float>(blk, na, nb, ri, ci, out, ld, mirror, tmb::as_stream(stream));
}
int tm_scatter_block_f64(const double* blk, int64_t na, int64_t nb, const int64_t* ri,
                         const int64_t* ci, double* out, int64_t ld, int mirror,
                         tm_stream_t stream) {
    return tmb::scatter_block<double>(blk, na, nb, ri, ci, out, ld, mirror, tmb::as_stream(stream));
}
int tm_scatter_diag_f32(const float* diag, int64_t na, const int64_t* ri, double* out, int64_t ld,
                        tm_stream_t stream) {
    return tmb::scatter_diag<float>(diag, na, ri, out, ld, tmb::as_stream(stream));
}
int tm_scatter_diag_f64(const double* diag, int64_t na, const int64_t* ri, double* out, int64_t ld,
                        tm_stream_t stream) {
    return tmb::scatter_diag<double>(diag, na, ri, out, ld, tmb::as_stream(stream));
}

}  // extern "C"
