// Dense block: CUDA-core kernels (fp64, restricted fp32, and the fp32 fallback) for
// sandwich, matvec, rmatvec and the weighted column second moment.
// The fp32 unrestricted-column sandwich is served by the tcgen05 kernel in dense_tc.cu.
//
// Reference semantics: dense.pyx:19-122, dense_helpers-tmpl.cpp:161-417.
#include <cstdlib>

#include "tm_common.cuh"

namespace tmb {

// ---------------------------------------------------------------------------------------
// Generic weighted SYRK:  out[a,b] = sum_t X[rows[t],cols[a]] * d[rows[t]] * X[rows[t],cols[b]]
// 64x64 output tiles (lower-triangular tile pairs only, like the reference's Ci>=Cj loop,
// dense_helpers-tmpl.cpp:236), split-K over row chunks, 4x4 register micro-tile,
// d folded in while staging the B tile in shared memory (the reference folds d while
// packing R, dense_helpers-tmpl.cpp:224,229).  Partial tiles are accumulated with RED.
// ---------------------------------------------------------------------------------------
constexpr int GS_BM = 64;
constexpr int GS_BK = 16;

template <typename F, bool C_ORDER>
__global__ void __launch_bounds__(256)
k_dense_sandwich_generic(const F* __restrict__ X, int64_t n, int64_t p,
                         const F* __restrict__ d, const int32_t* __restrict__ rows,
                         int64_t n_rows, const int32_t* __restrict__ cols, int64_t m,
                         F* __restrict__ out, int64_t rows_per_split) {
    __shared__ F As[GS_BK][GS_BM + 4];
    __shared__ F Bs[GS_BK][GS_BM + 4];

    // decode lower-triangular tile pair
    int idx = blockIdx.x;
    int ti = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= idx) ++ti;
    while (ti * (ti + 1) / 2 > idx) --ti;
    int tj = idx - ti * (ti + 1) / 2;

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t a0 = (int64_t)ti * GS_BM, b0 = (int64_t)tj * GS_BM;

    int64_t t_begin = (int64_t)blockIdx.y * rows_per_split;
    int64_t t_end = t_begin + rows_per_split;
    if (t_end > n_rows) t_end = n_rows;

    F acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = F(0);

    for (int64_t t0 = t_begin; t0 < t_end; t0 += GS_BK) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            int e = tid + l * 256;
            int kk, cc;
            if (C_ORDER) {
                kk = e >> 6;
                cc = e & 63;
            } else {
                kk = e & 15;
                cc = e >> 4;
            }
            int64_t t = t0 + kk;
            F av = F(0), bv = F(0);
            if (t < t_end) {
                int64_t k = row_at(rows, t);
                F dk = d[k];
                int64_t ca = a0 + cc, cb = b0 + cc;
                if (ca < m) {
                    int64_t j = cols ? (int64_t)cols[ca] : ca;
                    av = C_ORDER ? X[k * p + j] : X[j * n + k];
                }
                if (cb < m) {
                    int64_t j = cols ? (int64_t)cols[cb] : cb;
                    bv = (C_ORDER ? X[k * p + j] : X[j * n + k]) * dk;
                }
            }
            As[kk][cc] = av;
            Bs[kk][cc] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GS_BK; ++kk) {
            F a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t a = a0 + ty * 4 + i;
        if (a >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t b = b0 + tx * 4 + j;
            if (b >= m || b > a) continue;  // strict upper part is mirrored later
            red_add(&out[a * m + b], acc[i][j]);
        }
    }
}

template <typename F>
int dense_sandwich_generic(const F* X, int64_t n, int64_t p, int c_order, const F* d,
                           const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t m,
                           F* out, cudaStream_t st) {
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)(m * m), st));
    if (n_rows <= 0 || m <= 0) return 0;
    int64_t T = (m + GS_BM - 1) / GS_BM;
    int64_t npairs = T * (T + 1) / 2;
    int64_t want = (int64_t)sm_count() * 4;
    int64_t ksplit = (want + npairs - 1) / npairs;
    int64_t max_split = (n_rows + GS_BK * 16 - 1) / (GS_BK * 16);
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 65535) ksplit = 65535;
    int64_t rps = (n_rows + ksplit - 1) / ksplit;
    rps = (rps + GS_BK - 1) / GS_BK * GS_BK;
    ksplit = (n_rows + rps - 1) / rps;
    dim3 grid((unsigned)npairs, (unsigned)ksplit);
    if (c_order)
        k_dense_sandwich_generic<F, true><<<grid, 256, 0, st>>>(X, n, p, d, rows, n_rows, cols, m,
                                                                 out, rps);
    else
        k_dense_sandwich_generic<F, false><<<grid, 256, 0, st>>>(X, n, p, d, rows, n_rows, cols, m,
                                                                  out, rps);
    TM_LAUNCHED();
    return symmetrize_from_lower<F>(out, m, st);
}

// ---------------------------------------------------------------------------------------
// fp64 weighted SYRK on the FP64 tensor path: mma.sync m8n8k4 f64 (DMMA; tcgen05 has no fp64
// kind).  128x128 output tiles (lower-triangular tile pairs), 8 warps as 4 (M) x 2 (N), warp
// tile 32 x 64 = 4 x 8 m8n8 fragments, split-K over row chunks, d folded while staging the
// B tile, fragments above the diagonal of a diagonal tile skipped.  Same restrictions
// (rows / cols index lists, C or F order) as the generic kernel.
// ---------------------------------------------------------------------------------------
constexpr int DM_T = 128;   // output tile edge
constexpr int DM_BK = 16;   // rows per shared-memory stage
constexpr int DM_LD = 132;  // padded leading dimension (conflict-free fragment loads)

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool C_ORDER>
__global__ void __launch_bounds__(256)
k_dense_sandwich_dmma(const double* __restrict__ X, int64_t n, int64_t p,
                      const double* __restrict__ d, const int32_t* __restrict__ rows,
                      int64_t n_rows, const int32_t* __restrict__ cols, int64_t m,
                      double* __restrict__ out, int64_t rows_per_split) {
    __shared__ double As[DM_BK][DM_LD];
    __shared__ double Bs[DM_BK][DM_LD];

    int idx = blockIdx.x;
    int ti = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= idx) ++ti;
    while (ti * (ti + 1) / 2 > idx) --ti;
    const int tj = idx - ti * (ti + 1) / 2;
    const bool diag_tile = ti == tj;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;  // warp tile: rows wm*32.., cols wn*64..
    const int64_t a0 = (int64_t)ti * DM_T, b0 = (int64_t)tj * DM_T;
    int64_t t_begin = (int64_t)blockIdx.y * rows_per_split;
    int64_t t_end = t_begin + rows_per_split;
    if (t_end > n_rows) t_end = n_rows;

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // fragments of this warp that lie entirely above the diagonal of a diagonal tile are skipped
    unsigned keep = 0;  // bit (i*8+j)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int rmax = wm * 32 + i * 8 + 7, cmin = wn * 64 + j * 8;
            if (!diag_tile || rmax >= cmin) keep |= 1u << (i * 8 + j);
        }

    for (int64_t t0 = t_begin; t0 < t_end; t0 += DM_BK) {
#pragma unroll
        for (int l = 0; l < (DM_BK * DM_T) / 256; ++l) {
            const int e = tid + l * 256;
            int kk, cc;
            if (C_ORDER) {
                kk = e >> 7;
                cc = e & 127;
            } else {
                kk = e & (DM_BK - 1);
                cc = e >> 4;
            }
            const int64_t t = t0 + kk;
            double av = 0.0, bv = 0.0;
            if (t < t_end) {
                const int64_t k = row_at(rows, t);
                const double dk = d[k];
                const int64_t ca = a0 + cc, cb = b0 + cc;
                if (ca < m) {
                    const int64_t j = cols ? (int64_t)cols[ca] : ca;
                    av = C_ORDER ? X[k * p + j] : X[j * n + k];
                }
                if (diag_tile) {
                    bv = av * dk;
                } else if (cb < m) {
                    const int64_t j = cols ? (int64_t)cols[cb] : cb;
                    bv = (C_ORDER ? X[k * p + j] : X[j * n + k]) * dk;
                }
            }
            As[kk][cc] = av;
            Bs[kk][cc] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < DM_BK / 4; ++ks) {
            const int kr = ks * 4 + (lane & 3);
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kr][wm * 32 + i * 8 + (lane >> 2)];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = Bs[kr][wn * 64 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (keep & (1u << (i * 8 + j))) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t a = a0 + wm * 32 + i * 8 + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (!(keep & (1u << (i * 8 + j)))) continue;
            const int64_t b = b0 + wn * 64 + j * 8 + (lane & 3) * 2;
            if (a < m) {
                if (b < m && b <= a) red_add(&out[a * m + b], acc[i][j][0]);
                if (b + 1 < m && b + 1 <= a) red_add(&out[a * m + b + 1], acc[i][j][1]);
            }
        }
    }
}

int dense_sandwich_dmma(const double* X, int64_t n, int64_t p, int c_order, const double* d,
                        const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t m,
                        double* out, cudaStream_t st) {
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)(m * m), st));
    if (n_rows <= 0 || m <= 0) return 0;
    int64_t T = (m + DM_T - 1) / DM_T;
    int64_t npairs = T * (T + 1) / 2;
    int64_t want = (int64_t)sm_count() * 2;
    int64_t ksplit = (want + npairs - 1) / npairs;
    int64_t max_split = (n_rows + DM_BK * 8 - 1) / (DM_BK * 8);
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 65535) ksplit = 65535;
    int64_t rps = (n_rows + ksplit - 1) / ksplit;
    rps = (rps + DM_BK - 1) / DM_BK * DM_BK;
    ksplit = (n_rows + rps - 1) / rps;
    dim3 grid((unsigned)npairs, (unsigned)ksplit);
    if (c_order)
        k_dense_sandwich_dmma<true><<<grid, 256, 0, st>>>(X, n, p, d, rows, n_rows, cols, m, out, rps);
    else
        k_dense_sandwich_dmma<false><<<grid, 256, 0, st>>>(X, n, p, d, rows, n_rows, cols, m, out, rps);
    TM_LAUNCHED();
    return symmetrize_from_lower<double>(out, m, st);
}

// ---------------------------------------------------------------------------------------
// matvec:  out[r] (+)= sum_c X[rows[r], cols[c]] * v[cols[c]]
// ---------------------------------------------------------------------------------------
template <typename F>
__global__ void k_dense_matvec_c(const F* __restrict__ X, int64_t p, const F* __restrict__ v,
                                 const int32_t* __restrict__ rows, int64_t n_rows,
                                 const int32_t* __restrict__ cols, int64_t n_cols,
                                 F* __restrict__ out, int accumulate) {
    // one warp per row (C order: a row is contiguous)
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += nwarps) {
        int64_t k = row_at(rows, r);
        const F* xr = X + k * p;
        F s = F(0);
        for (int64_t c = lane; c < n_cols; c += 32) {
            int64_t j = cols ? (int64_t)cols[c] : c;
            s = fma(xr[j], v[j], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[r] = accumulate ? out[r] + s : s;
    }
}

template <typename F>
__global__ void k_dense_matvec_f(const F* __restrict__ X, int64_t n, const F* __restrict__ v,
                                 const int32_t* __restrict__ rows, int64_t n_rows,
                                 const int32_t* __restrict__ cols, int64_t n_cols,
                                 F* __restrict__ out, int accumulate) {
    // one thread per row (F order: consecutive rows are contiguous inside a column)
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        int64_t k = row_at(rows, r);
        F s = F(0);
        for (int64_t c = 0; c < n_cols; ++c) {
            int64_t j = cols ? (int64_t)cols[c] : c;
            s = fma(X[j * n + k], v[j], s);
        }
        out[r] = accumulate ? out[r] + s : s;
    }
}

template <typename F>
int dense_matvec(const F* X, int64_t n, int64_t p, int c_order, const F* v, const int32_t* rows,
                 int64_t n_rows, const int32_t* cols, int64_t n_cols, F* out, int accumulate,
                 cudaStream_t st) {
    if (n_rows <= 0) return 0;
    if (c_order) {
        int g = grid_for(n_rows * 32, 256, sm_count() * 32);
        k_dense_matvec_c<F><<<g, 256, 0, st>>>(X, p, v, rows, n_rows, cols, n_cols, out, accumulate);
    } else {
        int g = grid_for(n_rows, 256, sm_count() * 32);
        k_dense_matvec_f<F><<<g, 256, 0, st>>>(X, n, v, rows, n_rows, cols, n_cols, out, accumulate);
    }
    TM_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// column reductions over rows:
//   MODE 0 (rmatvec):        out[c] = sum_t X[rows[t], cols[c]] * w[rows[t]]
//   MODE 1 (sq_dot_weights): out[c] = sum_t w[rows[t]] * (X[rows[t], cols[c]] - shift[cols[c]])^2
// ---------------------------------------------------------------------------------------
template <typename F, int MODE>
__device__ __forceinline__ F col_term(F x, F w, F shift) {
    if (MODE == 0) return x * w;
    F t = x - shift;
    return w * t * t;
}

template <typename F, int MODE>
__global__ void __launch_bounds__(256)
k_dense_colreduce_c(const F* __restrict__ X, int64_t p, const F* __restrict__ w,
                    const F* __restrict__ shift, const int32_t* __restrict__ rows,
                    int64_t n_rows, const int32_t* __restrict__ cols, int64_t n_cols,
                    F* __restrict__ out, int64_t rows_per_block) {
    // block = 32 (columns) x 8 (row lanes); grid.x = row chunks, grid.y = column groups of 32
    __shared__ F red[8][33];
    int64_t c = (int64_t)blockIdx.y * 32 + threadIdx.x;
    int64_t t0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t t1 = t0 + rows_per_block;
    if (t1 > n_rows) t1 = n_rows;
    F s = F(0);
    if (c < n_cols) {
        int64_t j = cols ? (int64_t)cols[c] : c;
        F sh = (MODE == 1) ? shift[j] : F(0);
        for (int64_t t = t0 + threadIdx.y; t < t1; t += 8) {
            int64_t k = row_at(rows, t);
            s += col_term<F, MODE>(X[k * p + j], w[k], sh);
        }
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < n_cols) {
        F tot = F(0);
#pragma unroll
        for (int y = 0; y < 8; ++y) tot += red[y][threadIdx.x];
        red_add(&out[c], tot);
    }
}

template <typename F, int MODE>
__global__ void __launch_bounds__(256)
k_dense_colreduce_f(const F* __restrict__ X, int64_t n, const F* __restrict__ w,
                    const F* __restrict__ shift, const int32_t* __restrict__ rows,
                    int64_t n_rows, const int32_t* __restrict__ cols, int64_t n_cols,
                    F* __restrict__ out, int64_t rows_per_block) {
    // grid.x = row chunks, grid.y = column; whole block reduces one column chunk
    __shared__ F red[8];
    int64_t c = blockIdx.y;
    int64_t j = cols ? (int64_t)cols[c] : c;
    F sh = (MODE == 1) ? shift[j] : F(0);
    int64_t t0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t t1 = t0 + rows_per_block;
    if (t1 > n_rows) t1 = n_rows;
    F s = F(0);
    const F* xc = X + j * n;
    for (int64_t t = t0 + threadIdx.x; t < t1; t += 256) {
        int64_t k = row_at(rows, t);
        s += col_term<F, MODE>(xc[k], w[k], sh);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        F tot = F(0);
#pragma unroll
        for (int y = 0; y < 8; ++y) tot += red[y];
        red_add(&out[c], tot);
    }
}

template <typename F, int MODE>
int dense_colreduce(const F* X, int64_t n, int64_t p, int c_order, const F* w, const F* shift,
                    const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                    F* out, cudaStream_t st) {
    if (n_cols <= 0) return 0;
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)n_cols, st));
    if (n_rows <= 0) return 0;
    if (c_order) {
        int64_t colgroups = (n_cols + 31) / 32;
        int64_t want = ((int64_t)sm_count() * 8 + colgroups - 1) / colgroups;
        int64_t chunks = (n_rows + 255) / 256;
        if (chunks > want) chunks = want;
        if (chunks < 1) chunks = 1;
        int64_t rpb = (n_rows + chunks - 1) / chunks;
        chunks = (n_rows + rpb - 1) / rpb;
        dim3 grid((unsigned)chunks, (unsigned)colgroups), block(32, 8);
        k_dense_colreduce_c<F, MODE><<<grid, block, 0, st>>>(X, p, w, shift, rows, n_rows, cols,
                                                            n_cols, out, rpb);
    } else {
        int64_t want = ((int64_t)sm_count() * 8 + n_cols - 1) / n_cols;
        int64_t chunks = (n_rows + 2047) / 2048;
        if (chunks > want) chunks = want;
        if (chunks < 1) chunks = 1;
        int64_t rpb = (n_rows + chunks - 1) / chunks;
        chunks = (n_rows + rpb - 1) / rpb;
        if (n_cols > 65535) return fail("dense_colreduce: more than 65535 columns in F order");
        dim3 grid((unsigned)chunks, (unsigned)n_cols);
        k_dense_colreduce_f<F, MODE><<<grid, 256, 0, st>>>(X, n, w, shift, rows, n_rows, cols,
                                                          n_cols, out, rpb);
    }
    TM_LAUNCHED();
    return 0;
}


// out[a * m + b] = full[cols[a] * p + cols[b]]
template <typename F>
__global__ void k_select_square(const F* __restrict__ full, int64_t p,
                                const int32_t* __restrict__ cols, int64_t m, F* __restrict__ out) {
    const int64_t total = m * m;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t a = i / m, b = i - a * m;
        out[i] = full[(int64_t)cols[a] * p + cols[b]];
    }
}

}  // namespace tmb

extern "C" {

int tm_dense_sandwich_f32(const float* X, int64_t n, int64_t p, int c_order, const float* d,
                          const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                          float* out, tm_stream_t stream) {
    cudaStream_t st = tmb::as_stream(stream);
    if (!rows) n_rows = n;
    if (!cols) n_cols = p;
    if (n_cols <= 0) return 0;
    // A column selection (dense.pyx:19-44 `cols`) stays on the tensor path while it keeps at
    // least a fifth of the columns: the kernel is bound by streaming X, so the full p x p costs
    // the same as any subset and a tiny kernel picks the selected rows / columns out of it.
    // Narrower selections are cheaper on the CUDA-core kernel (cost ~ n m^2).
    const bool cols_ok = cols == nullptr || n_cols * 5 >= p || tmb::g_dense_f32_mode == 2;
    bool want_tc = tmb::g_dense_f32_mode != 1 && cols_ok && n_rows > 0 &&
                   tmb::dense_tc_eligible(n, p, c_order, X);
    if (tmb::g_dense_f32_mode == 2 && !want_tc)
        return tmb::fail("tm_dense_sandwich_f32: tcgen05 path forced but not eligible");
    if (want_tc) {
        tmb::Scratch dm(rows ? sizeof(float) * (size_t)n : 0, st);
        tmb::Scratch full(cols ? sizeof(float) * (size_t)(p * p) : 0, st);
        if (dm.err != cudaSuccess) return tmb::fail_cuda(dm.err, "scratch");
        if (full.err != cudaSuccess) return tmb::fail_cuda(full.err, "scratch");
        if (rows) {
            // fold the row restriction into the weights
            int rc = tmb::masked_weights<float>(d, n, rows, n_rows, dm.as<float>(), st);
            if (rc) return rc;
            d = dm.as<float>();
        }
        int rc = tmb::dense_sandwich_tc_f32(X, n, p, c_order, d, cols ? full.as<float>() : out, st);
        if (rc || !cols) return rc;
        tmb::k_select_square<float><<<tmb::grid_for(n_cols * n_cols, 256, tmb::sm_count() * 8), 256,
                                      0, st>>>(full.as<float>(), p, cols, n_cols, out);
        TM_LAUNCHED();
        return 0;
    }
    return tmb::dense_sandwich_generic<float>(X, n, p, c_order, d, rows, n_rows, cols, n_cols, out,
                                             st);
}

int tm_dense_sandwich_f64(const double* X, int64_t n, int64_t p, int c_order, const double* d,
                          const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t n_cols,
                          double* out, tm_stream_t stream) {
    if (!rows) n_rows = n;
    if (!cols) n_cols = p;
    if (n_cols <= 0) return 0;
    // fp64: DMMA tensor path (TABMAT_B200_F64_GENERIC=1 selects the CUDA-core kernel)
    static const bool f64_generic =
        getenv("TABMAT_B200_F64_GENERIC") && atoi(getenv("TABMAT_B200_F64_GENERIC")) == 1;
    if (f64_generic)
        return tmb::dense_sandwich_generic<double>(X, n, p, c_order, d, rows, n_rows, cols, n_cols,
                                                  out, tmb::as_stream(stream));
    return tmb::dense_sandwich_dmma(X, n, p, c_order, d, rows, n_rows, cols, n_cols, out,
                                    tmb::as_stream(stream));
}

#define TM_DENSE_VEC_API(SUF, F)                                                                 \
    int tm_dense_matvec_##SUF(const F* X, int64_t n, int64_t p, int c_order, const F* v,         \
                              const int32_t* rows, int64_t n_rows, const int32_t* cols,          \
                              int64_t n_cols, F* out, int accumulate, tm_stream_t stream) {      \
        if (!rows) n_rows = n;                                                                   \
        if (!cols) n_cols = p;                                                                   \
        return tmb::dense_matvec<F>(X, n, p, c_order, v, rows, n_rows, cols, n_cols, out,         \
                                   accumulate, tmb::as_stream(stream));                           \
    }                                                                                            \
    int tm_dense_rmatvec_##SUF(const F* X, int64_t n, int64_t p, int c_order, const F* v,        \
                               const int32_t* rows, int64_t n_rows, const int32_t* cols,         \
                               int64_t n_cols, F* out, tm_stream_t stream) {                     \
        if (!rows) n_rows = n;                                                                   \
        if (!cols) n_cols = p;                                                                   \
        return tmb::dense_colreduce<F, 0>(X, n, p, c_order, v, nullptr, rows, n_rows, cols,       \
                                         n_cols, out, tmb::as_stream(stream));                    \
    }                                                                                            \
    int tm_dense_sq_dot_weights_##SUF(const F* X, int64_t n, int64_t p, int c_order, const F* w, \
                                      const F* shift, F* out, tm_stream_t stream) {              \
        return tmb::dense_colreduce<F, 1>(X, n, p, c_order, w, shift, nullptr, n, nullptr, p,     \
                                         out, tmb::as_stream(stream));                            \
    }

TM_DENSE_VEC_API(f32, float)
TM_DENSE_VEC_API(f64, double)

}  // extern "C"
