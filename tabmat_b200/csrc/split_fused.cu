// Fused cross blocks of a SplitMatrix sandwich that share the dense operand.
//
// The reference computes dense x categorical_i (categorical_matrix.py:759-791 ->
// split.pyx:32-80) and dense x sparse (sparse_matrix.py:206-229 -> sparse.pyx:211-260) with one
// native call per pair inside the Python block loop of split_matrix.py:346-354, re-reading the
// dense block every time.  Here ONE pass streams every dense row once (one coalesced 16-byte
// load per lane), scales it by d[k], and adds it to
//     out_cat_i[codes_i[k] - drop_first_i, :]        for every categorical block i
//     out_sparse[j, :] * A[k, j]                      for every non-zero (k, j) of the sparse block
// with vector RED.ADD into L2.  Measured on B200: the L2 atomic units sustain ~6.0 TB/s of RED
// payload when the destination rows are spread over >= ~1300 distinct rows, and collapse under
// contention (0.6 TB/s on 10 rows), so categorical blocks with few levels are accumulated into
// `copies` replicas of their table (replica = warp id mod copies) that a second tiny kernel sums.
#include <cstdlib>

#include "tm_common.cuh"

namespace tmb {

__device__ __forceinline__ void red_add_vec(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add_vec(double* p, double2 v) {
    atomicAdd(p, v.x);
    atomicAdd(p + 1, v.y);
}

template <typename F>
struct Vec;
template <>
struct Vec<float> {
    using T = float4;
    static constexpr int W = 4;
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ T scale(T a, float s) {
        return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
    }
    static __device__ __forceinline__ T add(T a, T b) {
        return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
};
template <>
struct Vec<double> {
    using T = double2;
    static constexpr int W = 2;
    static __device__ __forceinline__ T zero() { return make_double2(0.0, 0.0); }
    static __device__ __forceinline__ T scale(T a, double s) { return make_double2(a.x * s, a.y * s); }
    static __device__ __forceinline__ T add(T a, T b) { return make_double2(a.x + b.x, a.y + b.y); }
};

__device__ __forceinline__ const void* shfl_ptr(const void* p, int src) {
    unsigned long long v = reinterpret_cast<unsigned long long>(p);
    unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return reinterpret_cast<const void*>(((unsigned long long)hi << 32) | lo);
}

// One warp per row; lane l owns the column chunks l, l+32, ... (W columns each); NV chunks/lane.
// Destination rows are computed lane-parallel (lane i: categorical block i / the i-th non-zero
// of the row) and broadcast with shuffles, so the per-destination cost is two shuffles, one
// 64-bit add and the vector RED.
template <typename F, int NV>
__global__ void __launch_bounds__(256)
k_dense_cross_fused(const F* __restrict__ X, int64_t n, int P, const F* __restrict__ d,
                    const int32_t* __restrict__ rows, int64_t n_rows,
                    const FusedCrossParams prm) {
    using V = Vec<F>;
    using VT = typename V::T;
    constexpr int W = V::W;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int chunks = P / W;  // P % W == 0 (checked by the host)
    const F* csr_data = static_cast<const F*>(prm.csr_data);
    F* out_sparse = static_cast<F*>(prm.out_sparse);
    const int n_cat = prm.n_cat;

    // lane i < n_cat serves categorical block i: its code vector, drop_first and the base of
    // the table replica this warp adds into
    const int32_t* my_codes = nullptr;
    int my_df = 0;
    F* my_tab = nullptr;
    if (lane < n_cat) {
        my_codes = prm.codes[lane];
        my_df = prm.drop_first[lane];
        my_tab = static_cast<F*>(prm.tab[lane]) +
                 (int64_t)(warp % prm.copies[lane]) * prm.K[lane] * (int64_t)P;
    }
    const int64_t lane_off = (int64_t)lane * W;

    constexpr int U = 2;  // rows in flight per warp: all loads of U rows are issued before the REDs
    for (int64_t t = warp; t < n_rows; t += (int64_t)U * nwarps) {
        F dk[U];
        VT y[U][NV];
        const F* cat_dst[U];   // lane i: destination row of categorical block i (or nullptr)
        int e0[U], e1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t tu = t + (int64_t)u * nwarps;
            const bool ok = tu < n_rows;
            const int64_t k = ok ? row_at(rows, tu) : 0;
            dk[u] = ok ? d[k] : F(0);
            const VT* xr = reinterpret_cast<const VT*>(X + k * (int64_t)P);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                y[u][v] = (ok && c < chunks) ? __ldg(xr + c) : V::zero();
            }
            cat_dst[u] = nullptr;
            if (ok && lane < n_cat) {
                const int c = my_codes[k] - my_df;
                if (c >= 0) cat_dst[u] = my_tab + (int64_t)c * P;
            }
            e0[u] = e1[u] = 0;
            if (ok && out_sparse) {
                e0[u] = prm.csr_indptr[k];
                e1[u] = prm.csr_indptr[k + 1];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (dk[u] == F(0)) continue;  // every term of row k is proportional to d[k]
#pragma unroll
            for (int v = 0; v < NV; ++v) y[u][v] = V::scale(y[u][v], dk[u]);
            for (int i = 0; i < n_cat; ++i) {
                F* dst = const_cast<F*>(static_cast<const F*>(shfl_ptr(cat_dst[u], i)));
                if (!dst) continue;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int ch = lane + 32 * v;
                    if (ch < chunks) red_add_vec(dst + lane_off + (int64_t)v * 32 * W, y[u][v]);
                }
            }
            for (int eb = e0[u]; eb < e1[u]; eb += 32) {
                const int e = eb + lane;
                const F* sp_dst = nullptr;
                F a = F(0);
                if (e < e1[u]) {
                    sp_dst = out_sparse + (int64_t)prm.csr_indices[e] * P;
                    a = csr_data[e];
                }
                const int cnt = min(32, e1[u] - eb);
                for (int q = 0; q < cnt; ++q) {
                    F* dst = const_cast<F*>(static_cast<const F*>(shfl_ptr(sp_dst, q)));
                    const F aa = __shfl_sync(0xffffffffu, a, q);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int ch = lane + 32 * v;
                        if (ch < chunks)
                            red_add_vec(dst + lane_off + (int64_t)v * 32 * W, V::scale(y[u][v], aa));
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Run-aggregating variant for row-sorted SplitMatrices (tabmat_b200/row_order.py): a warp walks
// `chunk` CONSECUTIVE rows; for each categorical block it keeps d[k] * X[k, :] summed in
// registers for as long as the code does not change and issues ONE vector RED per run (and one
// at the end of the chunk).  When the rows are stored sorted by (code_a, code_b) the runs of
// block a span whole chunks and those of block b ~n / (K_a K_b) rows, so the categorical part
// of the L2-atomic payload (one 512-byte RED per row and block in k_dense_cross_fused) all but
// disappears; on unsorted rows every run has length 1 and the RED count is unchanged.  The
// sparse part is the same per-non-zero vector RED as above.
// Per 32-row group the row ids, weights, codes and CSR row bounds are loaded lane-parallel
// (coalesced) and broadcast with shuffles.
// ---------------------------------------------------------------------------------------
template <typename F, int NV, int NC, bool ILP>
__global__ void __launch_bounds__(256)
k_dense_cross_runs(const F* __restrict__ X, int64_t n, int P, const F* __restrict__ d,
                   const int32_t* __restrict__ rows, int64_t n_rows, int chunk,
                   const FusedCrossParams prm) {
    using V = Vec<F>;
    using VT = typename V::T;
    constexpr int W = V::W;
    constexpr int NCA = NC > 0 ? NC : 1;
    constexpr int U = 2;  // rows whose X loads are in flight together
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int chunks = P / W;
    const F* csr_data = static_cast<const F*>(prm.csr_data);
    F* out_sparse = static_cast<F*>(prm.out_sparse);
    const int64_t lane_off = (int64_t)lane * W;

    F* tab[NCA];
    const int32_t* codes[NCA];
    int df[NCA];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        codes[c] = prm.codes[c];
        df[c] = prm.drop_first[c];
        tab[c] = static_cast<F*>(prm.tab[c]) +
                 (int64_t)(warp % prm.copies[c]) * prm.K[c] * (int64_t)P;
    }

    const int64_t n_chunks = (n_rows + chunk - 1) / chunk;
    for (int64_t ci = warp; ci < n_chunks; ci += nwarps) {
        const int64_t t0 = ci * chunk;
        const int64_t t1 = t0 + chunk < n_rows ? t0 + chunk : n_rows;
        VT acc[NCA][NV];
        int cur[NCA];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            cur[c] = -1;
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[c][v] = V::zero();
        }
        for (int64_t tb = t0; tb < t1; tb += 32) {
            const int cnt = (int)(t1 - tb < 32 ? t1 - tb : 32);
            // lane-parallel metadata of the 32-row group
            int64_t k_l = 0;
            F d_l = F(0);
            int code_l[NCA];
            int e0_l = 0, e1_l = 0;
#pragma unroll
            for (int c = 0; c < NC; ++c) code_l[c] = -1;
            if (lane < cnt) {
                k_l = row_at(rows, tb + lane);
                d_l = d[k_l];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int cc = codes[c][k_l] - df[c];
                    code_l[c] = cc < 0 ? -1 : cc;
                }
                if (out_sparse) {
                    e0_l = prm.csr_indptr[k_l];
                    e1_l = prm.csr_indptr[k_l + 1];
                }
            }
            const unsigned klo_l = (unsigned)(unsigned long long)k_l;
            const unsigned khi_l = (unsigned)((unsigned long long)k_l >> 32);
            for (int q0 = 0; q0 < cnt; q0 += U) {
                VT x[U][NV];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int qq = q0 + u < cnt ? q0 + u : cnt - 1;
                    const unsigned lo = __shfl_sync(FULL, klo_l, qq);
                    const unsigned hi = __shfl_sync(FULL, khi_l, qq);
                    const int64_t k = (int64_t)(((unsigned long long)hi << 32) | lo);
                    const VT* xr = reinterpret_cast<const VT*>(X + k * (int64_t)P);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int ch = lane + 32 * v;
                        x[u][v] = ch < chunks ? __ldg(xr + ch) : V::zero();
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int q = q0 + u;
                    if (q >= cnt) break;
                    const F dk = __shfl_sync(FULL, d_l, q);
                    int code_q[NCA];
#pragma unroll
                    for (int c = 0; c < NC; ++c) code_q[c] = __shfl_sync(FULL, code_l[c], q);
                    const int e0 = __shfl_sync(FULL, e0_l, q);
                    const int e1 = __shfl_sync(FULL, e1_l, q);
                    if (dk == F(0)) continue;  // every term of row k is proportional to d[k]
                    VT y[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) y[v] = V::scale(x[u][v], dk);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        if (code_q[c] != cur[c]) {  // warp-uniform: a run of block c ends here
                            if (cur[c] >= 0) {
                                F* dst = tab[c] + (int64_t)cur[c] * P + lane_off;
#pragma unroll
                                for (int v = 0; v < NV; ++v) {
                                    if (lane + 32 * v < chunks)
                                        red_add_vec(dst + (int64_t)v * 32 * W, acc[c][v]);
                                    acc[c][v] = V::zero();
                                }
                            }
                            cur[c] = code_q[c];
                        }
                        if (code_q[c] >= 0) {
#pragma unroll
                            for (int v = 0; v < NV; ++v) acc[c][v] = V::add(acc[c][v], y[v]);
                        }
                    }
                    for (int eb = e0; eb < e1; eb += 32) {
                        const int e = eb + lane;
                        const F* sp_dst = nullptr;
                        F a = F(0);
                        if (e < e1) {
                            sp_dst = out_sparse + (int64_t)prm.csr_indices[e] * P;
                            a = csr_data[e];
                        }
                        const int m = min(32, e1 - eb);
                        // groups of 4 REDs with their own value registers: a RED holds its source
                        // registers until the LSU has taken the data, so reusing one register
                        // quadruple serialises a warp's REDs (one per ~500 cycles, B200)
                        if (!ILP) {
                            for (int z = 0; z < m; ++z) {
                                F* dst = const_cast<F*>(static_cast<const F*>(shfl_ptr(sp_dst, z)));
                                const F aa = __shfl_sync(FULL, a, z);
#pragma unroll
                                for (int v = 0; v < NV; ++v) {
                                    if (lane + 32 * v < chunks)
                                        red_add_vec(dst + lane_off + (int64_t)v * 32 * W,
                                                    V::scale(y[v], aa));
                                }
                            }
                        } else
                        for (int z0 = 0; z0 < m; z0 += 4) {
                            F* dst4[4];
                            VT val4[4][NV];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int z = (z0 + u) & 31;
                                dst4[u] = const_cast<F*>(static_cast<const F*>(shfl_ptr(sp_dst, z)));
                                const F aa = __shfl_sync(FULL, a, z);
#pragma unroll
                                for (int v = 0; v < NV; ++v) val4[u][v] = V::scale(y[v], aa);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (z0 + u >= m) break;
#pragma unroll
                                for (int v = 0; v < NV; ++v) {
                                    if (lane + 32 * v < chunks)
                                        red_add_vec(dst4[u] + lane_off + (int64_t)v * 32 * W,
                                                    val4[u][v]);
                                }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (cur[c] >= 0) {
                F* dst = tab[c] + (int64_t)cur[c] * P + lane_off;
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    if (lane + 32 * v < chunks) red_add_vec(dst + (int64_t)v * 32 * W, acc[c][v]);
            }
        }
    }
}

template <typename F, int NV>
static void launch_cross_runs(int nc, int g, cudaStream_t st, const F* X, int64_t n, int P,
                              const F* d, const int32_t* rows, int64_t n_rows, int chunk,
                              const FusedCrossParams& prm) {
    // TABMAT_B200_SCATTER_ILP=1: groups of 4 REDs with independent registers (80 registers per
    // thread -> 3 CTAs per SM instead of 8)
    static const bool ilp =
        getenv("TABMAT_B200_SCATTER_ILP") && atoi(getenv("TABMAT_B200_SCATTER_ILP")) == 1;
#define TM_RUNS(NCV)                                                                              \
    if (ilp)                                                                                      \
        k_dense_cross_runs<F, NV, NCV, true><<<g, 256, 0, st>>>(X, n, P, d, rows, n_rows, chunk, prm); \
    else                                                                                          \
        k_dense_cross_runs<F, NV, NCV, false><<<g, 256, 0, st>>>(X, n, P, d, rows, n_rows, chunk, prm);
    switch (nc) {
        case 0: TM_RUNS(0) break;
        case 1: TM_RUNS(1) break;
        case 2: TM_RUNS(2) break;
        case 3: TM_RUNS(3) break;
        default: TM_RUNS(4) break;
    }
#undef TM_RUNS
}

// ---------------------------------------------------------------------------------------
// categorical x dense by SORTED GATHER: `perm` lists the rows with a valid category ordered by
// category (built once per matrix, like the reference's cached CSR), `segptr[c]..segptr[c+1]`
// is category c's slice of it.  A warp walks a chunk of consecutive perm entries, gathers the
// 512-byte dense rows, accumulates d[k] * X[k, :] in registers and flushes once per category
// run - HBM-bound gather traffic instead of one L2 RED per row, so it can overlap with the
// RED-bound passes on another stream.
// ---------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
k_cat_dense_gather(const float* __restrict__ X, int P, const float* __restrict__ d,
                   const int32_t* __restrict__ perm, const int32_t* __restrict__ segptr, int K,
                   int64_t n_valid, int chunk, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int chunks4 = P / 4;
    const int64_t n_chunks = (n_valid + chunk - 1) / chunk;
    for (int64_t ci = warp; ci < n_chunks; ci += nwarps) {
        const int64_t t0 = ci * chunk;
        const int64_t t1 = t0 + chunk < n_valid ? t0 + chunk : n_valid;
        // category of entry t0: largest c with segptr[c] <= t0
        int lo = 0, hi = K;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if ((int64_t)segptr[mid] <= t0) lo = mid; else hi = mid;
        }
        int c = lo;
        int64_t bound = segptr[c + 1];
        float4 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        bool dirty = false;
        for (int64_t tb = t0; tb < t1; tb += 32) {
            const int cnt = (int)(t1 - tb < 32 ? t1 - tb : 32);
            int k_l = 0;
            float d_l = 0.f;
            if (lane < cnt) {
                k_l = perm[tb + lane];
                d_l = d[k_l];
            }
            for (int q0 = 0; q0 < cnt; q0 += 4) {
                float4 x[4][NV];
                float dk[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int qq = q0 + u < cnt ? q0 + u : cnt - 1;
                    const int k = __shfl_sync(0xffffffffu, k_l, qq);
                    dk[u] = q0 + u < cnt ? __shfl_sync(0xffffffffu, d_l, qq) : 0.f;
                    const float4* xr = reinterpret_cast<const float4*>(X + (int64_t)k * P);
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int ch = lane + 32 * v;
                        x[u][v] = ch < chunks4 ? __ldg(xr + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t t = tb + q0 + u;
                    if (t >= t1) break;
                    if (t >= bound) {  // category run ended: flush and advance (warp-uniform)
                        if (dirty) {
#pragma unroll
                            for (int v = 0; v < NV; ++v) {
                                const int ch = lane + 32 * v;
                                if (ch < chunks4) red_add_vec(out + (int64_t)c * P + ch * 4, acc[v]);
                                acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            dirty = false;
                        }
                        while (t >= bound) {
                            ++c;
                            bound = segptr[c + 1];
                        }
                    }
                    if (dk[u] != 0.f) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            acc[v].x = fmaf(dk[u], x[u][v].x, acc[v].x);
                            acc[v].y = fmaf(dk[u], x[u][v].y, acc[v].y);
                            acc[v].z = fmaf(dk[u], x[u][v].z, acc[v].z);
                            acc[v].w = fmaf(dk[u], x[u][v].w, acc[v].w);
                        }
                        dirty = true;
                    }
                }
            }
        }
        if (dirty) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int ch = lane + 32 * v;
                if (ch < chunks4) red_add_vec(out + (int64_t)c * P + ch * 4, acc[v]);
            }
        }
    }
}

// out (K x p) = sum over the sorted rows; overwrites out.  f32, row-major X, p % 4 == 0, p <= 256.
int cat_dense_gather_f32(const float* X, int64_t p, const float* d, const int32_t* perm,
                         const int32_t* segptr, int64_t K, int64_t n_valid, float* out,
                         cudaStream_t st) {
    if (K <= 0 || p <= 0) return 0;
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(K * p), st));
    if (n_valid <= 0) return 0;
    const int chunk = 256;
    int64_t n_chunks = (n_valid + chunk - 1) / chunk;
    int g = grid_for(n_chunks * 32, 256, sm_count() * 6);
    if (p <= 128)
        k_cat_dense_gather<1><<<g, 256, 0, st>>>(X, (int)p, d, perm, segptr, (int)K, n_valid, chunk, out);
    else
        k_cat_dense_gather<2><<<g, 256, 0, st>>>(X, (int)p, d, perm, segptr, (int)K, n_valid, chunk, out);
    TM_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------
// dense x sparse by ROW-BLOCKED GATHER (the alternative to one vector RED per non-zero):
//   out[j, :] = sum over the non-zeros (k, j) of column j of  A[k, j] * d[k] * X[k, :]
// The L2 atomic units cap the RED form at 1.9e11 sector-ops/s (16 per non-zero at 128 fp32
// columns: 11.5 ms at the benchmark shape); the L2 read path is about twice as wide.  Here the
// non-zeros are stored ordered by (row block, column, row) — a second row-blocked CSC copy,
// built once per matrix, with blocks of `block_rows` rows chosen so that a block of X
// (block_rows * P * sizeof F <= 32 MB) stays in the 126 MB L2.  A warp owns a fixed set of
// columns, walks the row blocks in order, gathers the 512-byte X rows of a (block, column) run
// with coalesced 16-byte loads, sums them in registers and issues ONE vector RED per run
// (~65 non-zeros at the benchmark shape: 65x fewer REDs).  All CTAs are resident and every warp
// has the same number of columns, so the whole grid sweeps the row blocks together; a soft
// barrier (per-block arrival counters, bounded spin: a hint, never needed for correctness)
// keeps any CTA from running more than `lag` blocks ahead, so the gather window stays in L2.
// ---------------------------------------------------------------------------------------
constexpr int GD_THREADS = 128;

template <typename F, int NV, int U>
__global__ void __launch_bounds__(GD_THREADS)
k_csc_dense_gather(const F* __restrict__ X, int P, const F* __restrict__ d,
                   const F* __restrict__ bdata, const int32_t* __restrict__ brow,
                   const int32_t* __restrict__ bptr, int p_s, int n_blocks, int cols_per_warp,
                   F* __restrict__ out, unsigned* __restrict__ progress, int lag) {
    using V = Vec<F>;
    using VT = typename V::T;
    constexpr int W = V::W;   // U = X rows in flight per warp
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gw = (int)(((int64_t)blockIdx.x * GD_THREADS + threadIdx.x) >> 5);
    const int chunks = P / W;
    const int j0 = gw * cols_per_warp;
    const int j1 = min(p_s, j0 + cols_per_warp);
    for (int b = 0; b < n_blocks; ++b) {
        if (b >= lag) {
            if (threadIdx.x == 0) {
                const volatile unsigned* pr = progress + (b - lag);
                for (int spin = 0; spin < 4096 && *pr < gridDim.x; ++spin) __nanosleep(64);
            }
            __syncthreads();
        }
        const int32_t* ptr = bptr + (int64_t)b * p_s;
        for (int j = j0; j < j1; ++j) {
            const int e0 = ptr[j], e1 = ptr[j + 1];
            if (e0 == e1) continue;
            VT acc[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = V::zero();
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                int k_l = 0;
                F w_l = F(0);
                if (e < e1) {
                    k_l = __ldg(brow + e);
                    w_l = __ldg(bdata + e) * __ldg(d + k_l);
                }
                const int cnt = min(32, e1 - eb);
                for (int z0 = 0; z0 < cnt; z0 += U) {
                    VT x[U][NV];
                    F w[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int z = (z0 + u) & 31;   // lanes past cnt carry weight 0, row 0
                        const int k = __shfl_sync(FULL, k_l, z);
                        w[u] = __shfl_sync(FULL, w_l, z);
                        const VT* xr = reinterpret_cast<const VT*>(X + (int64_t)k * P);
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const int ch = lane + 32 * v;
                            x[u][v] = ch < chunks ? __ldg(xr + ch) : V::zero();
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (z0 + u >= cnt) break;
#pragma unroll
                        for (int v = 0; v < NV; ++v) acc[v] = V::add(acc[v], V::scale(x[u][v], w[u]));
                    }
                }
            }
            F* dst = out + (int64_t)j * P + (int64_t)lane * W;
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (lane + 32 * v < chunks) red_add_vec(dst + (int64_t)v * 32 * W, acc[v]);
        }
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(progress + b, 1u);
    }
}

// out (p_s x p) is overwritten.  bdata / brow / bptr: row-blocked CSC (n_blocks * p_s + 1 offsets).
template <typename F>
int csc_dense_gather(const F* X, int64_t p, const F* d, const F* bdata, const int32_t* brow,
                     const int32_t* bptr, int64_t p_s, int64_t n_blocks, F* out, cudaStream_t st) {
    constexpr int W = Vec<F>::W;
    if (p <= 0 || p % W != 0 || p > 64 * W) return fail("csc_dense_gather: unsupported dense width");
    if (p_s <= 0 || n_blocks <= 0) return 0;
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(F) * (size_t)(p_s * p), st));
    Scratch prog(sizeof(unsigned) * (size_t)n_blocks, st);
    if (prog.err != cudaSuccess) return fail_cuda(prog.err, "scratch");
    TM_CUDA(cudaMemsetAsync(prog.p, 0, sizeof(unsigned) * (size_t)n_blocks, st));
    const int nv = (int)((p / W + 31) / 32);
    // rows in flight per warp (TABMAT_B200_GATHER_U: 8 default, 16) and CTAs per SM
    // (TABMAT_B200_GATHER_CTAS, default 8 = as many as fit): the kernel is bound by L2 latency x bytes in flight
    static const int u_rows = [] {
        const char* e = getenv("TABMAT_B200_GATHER_U");
        return (e && atoi(e) == 16) ? 16 : 8;
    }();
    static const int cta_cap = [] {
        const char* e = getenv("TABMAT_B200_GATHER_CTAS");
        int v = e ? atoi(e) : 8;   // measured: 14.6 ms at 4 (10 warps per SM), 8.0 ms at 8
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    // one resident wave, the same number of columns for every warp
    int per_sm = 0;
    if (nv == 1 && u_rows == 16)
        TM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_csc_dense_gather<F, 1, 16>,
                                                              GD_THREADS, 0));
    else if (nv == 1)
        TM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_csc_dense_gather<F, 1, 8>,
                                                              GD_THREADS, 0));
    else
        TM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_csc_dense_gather<F, 2, 8>,
                                                              GD_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > cta_cap) per_sm = cta_cap;
    // SMs left free for a concurrent collective (tm_set_sm_reserve; row-sharded callers overlap
    // the allreduce of the index blocks with this kernel)
    int sms = sm_count() - g_sm_reserve;
    if (sms < 8) sms = 8;
    const int64_t max_warps = (int64_t)sms * per_sm * (GD_THREADS / 32);
    const int cpw = (int)((p_s + max_warps - 1) / max_warps);
    const int64_t warps = (p_s + cpw - 1) / cpw;
    const int grid = (int)((warps + GD_THREADS / 32 - 1) / (GD_THREADS / 32));
    static const int lag = [] {
        const char* e = getenv("TABMAT_B200_GATHER_LAG");
        int v = e ? atoi(e) : 2;
        return v < 1 ? 1 : v;
    }();
    if (nv == 1 && u_rows == 16)
        k_csc_dense_gather<F, 1, 16><<<grid, GD_THREADS, 0, st>>>(
            X, (int)p, d, bdata, brow, bptr, (int)p_s, (int)n_blocks, cpw, out,
            prog.as<unsigned>(), lag);
    else if (nv == 1)
        k_csc_dense_gather<F, 1, 8><<<grid, GD_THREADS, 0, st>>>(
            X, (int)p, d, bdata, brow, bptr, (int)p_s, (int)n_blocks, cpw, out,
            prog.as<unsigned>(), lag);
    else
        k_csc_dense_gather<F, 2, 8><<<grid, GD_THREADS, 0, st>>>(
            X, (int)p, d, bdata, brow, bptr, (int)p_s, (int)n_blocks, cpw, out,
            prog.as<unsigned>(), lag);
    TM_LAUNCHED();
    return 0;
}
template int csc_dense_gather<float>(const float*, int64_t, const float*, const float*,
                                     const int32_t*, const int32_t*, int64_t, int64_t, float*,
                                     cudaStream_t);
template int csc_dense_gather<double>(const double*, int64_t, const double*, const double*,
                                      const int32_t*, const int32_t*, int64_t, int64_t, double*,
                                      cudaStream_t);

// out[r, :] = sum over replicas of tab[(rep*K + r), :]
template <typename F>
__global__ void k_sum_replicas(const F* __restrict__ tab, int K, int copies, int64_t P,
                               F* __restrict__ out) {
    int64_t total = (int64_t)K * P;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        F s = F(0);
        for (int r = 0; r < copies; ++r) s += tab[(int64_t)r * total + i];
        out[i] = s;
    }
}

template <typename F>
int cross_prepare(int64_t p, int n_cat, const int32_t* const* codes, const int64_t* K,
                  const int32_t* drop_first, F* const* out_cat, const F* csr_data,
                  const int32_t* csr_indices, const int32_t* csr_indptr, int64_t p_sparse,
                  F* out_sparse, FusedCrossParams& prm, CrossScratch& scr, cudaStream_t st) {
    if (n_cat > FC_MAX_CATS) return fail("tm_dense_cross_sandwich: more than 8 categorical blocks");
    memset(&prm, 0, sizeof(prm));
    prm.n_cat = n_cat;
    size_t scratch_elems = 0;
    int64_t offs[FC_MAX_CATS];
    for (int i = 0; i < n_cat; ++i) {
        prm.codes[i] = codes[i];
        prm.K[i] = (int)K[i];
        prm.drop_first[i] = drop_first[i];
        // spread hot tables over >= ~1536 destination rows (see header comment)
        int copies = 1;
        if (K[i] > 0 && K[i] < 1024) copies = (int)((1536 + K[i] - 1) / K[i]);
        prm.copies[i] = copies;
        offs[i] = -1;
        if (copies > 1) {
            offs[i] = (int64_t)scratch_elems;
            scratch_elems += (size_t)copies * (size_t)K[i] * (size_t)p;
        }
    }
    if (scratch_elems) {
        keep_pool_memory();
        TM_CUDA(cudaMallocAsync(&scr.p, scratch_elems * sizeof(F), st));
        scr.s = st;
        TM_CUDA(cudaMemsetAsync(scr.p, 0, scratch_elems * sizeof(F), st));
    }
    for (int i = 0; i < n_cat; ++i) {
        if (K[i] <= 0) continue;
        if (prm.copies[i] > 1) {
            prm.tab[i] = static_cast<F*>(scr.p) + offs[i];
        } else {
            prm.tab[i] = out_cat[i];
            TM_CUDA(cudaMemsetAsync(out_cat[i], 0, sizeof(F) * (size_t)(K[i] * p), st));
        }
    }
    if (out_sparse && p_sparse > 0) {
        prm.csr_data = csr_data;
        prm.csr_indices = csr_indices;
        prm.csr_indptr = csr_indptr;
        prm.out_sparse = out_sparse;
        TM_CUDA(cudaMemsetAsync(out_sparse, 0, sizeof(F) * (size_t)(p_sparse * p), st));
    }
    return 0;
}

template <typename F>
int cross_finish(int64_t p, int n_cat, const int64_t* K, F* const* out_cat,
                 const FusedCrossParams& prm, cudaStream_t st) {
    for (int i = 0; i < n_cat; ++i) {
        if (K[i] > 0 && prm.copies[i] > 1) {
            int g = grid_for(K[i] * p, 256, sm_count() * 4);
            k_sum_replicas<F><<<g, 256, 0, st>>>(static_cast<const F*>(prm.tab[i]), prm.K[i],
                                                 prm.copies[i], p, out_cat[i]);
            TM_LAUNCHED();
        }
    }
    return 0;
}
template int cross_prepare<float>(int64_t, int, const int32_t* const*, const int64_t*,
                                  const int32_t*, float* const*, const float*, const int32_t*,
                                  const int32_t*, int64_t, float*, FusedCrossParams&,
                                  CrossScratch&, cudaStream_t);
template int cross_finish<float>(int64_t, int, const int64_t*, float* const*,
                                 const FusedCrossParams&, cudaStream_t);

template <typename F>
int dense_cross_fused(const F* X, int64_t n, int64_t p, const F* d, const int32_t* rows,
                      int64_t n_rows, int n_cat, const int32_t* const* codes, const int64_t* K,
                      const int32_t* drop_first, F* const* out_cat, const F* csr_data,
                      const int32_t* csr_indices, const int32_t* csr_indptr, int64_t p_sparse,
                      F* out_sparse, int runs, cudaStream_t st) {
    constexpr int W = Vec<F>::W;
    // the run-aggregating kernel is also the faster one on unsorted rows (B200, n = 4e7:
    // 18.1 ms vs 19.7 ms), so it is the default; mode 2 selects the one-row-per-visit kernel
    if (g_cross_runs_mode == 1) runs = 1;
    if (g_cross_runs_mode == 2 || n_cat > 4) runs = 0;
    if (p <= 0 || p % W != 0 || p > 64 * W)
        return fail("tm_dense_cross_sandwich: unsupported dense width");
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0)
        return fail("tm_dense_cross_sandwich: X must be 16-byte aligned");
    if (!rows) n_rows = n;
    FusedCrossParams prm;
    CrossScratch scr;
    int rc = cross_prepare<F>(p, n_cat, codes, K, drop_first, out_cat, csr_data, csr_indices,
                              csr_indptr, p_sparse, out_sparse, prm, scr, st);
    if (rc) return rc;
    if (n_rows > 0 && runs) {
        // consecutive rows per warp and visit: long enough for the run aggregation to pay,
        // short enough that every warp of the grid gets several chunks
        const int64_t warps_total = (int64_t)sm_count() * 8 * 8;
        int64_t chunk = n_rows / (warps_total * 4);
        chunk = chunk < 32 ? 32 : (chunk > 128 ? 128 : chunk / 32 * 32);
        const int64_t n_chunks = (n_rows + chunk - 1) / chunk;
        // CTAs of 256 threads per SM: 8 fill every thread slot (64 warps keep ~64 REDs in flight
        // per SM); TABMAT_B200_SCATTER_CTAS lowers it so that a kernel of the index pass can be
        // co-resident (tm_split_sandwich_blocks schedules 2 / 4)
        static const int ctas = [] {
            const char* e = getenv("TABMAT_B200_SCATTER_CTAS");
            int v = e ? atoi(e) : 8;
            return v < 1 ? 1 : (v > 8 ? 8 : v);
        }();
        int g = grid_for(n_chunks * 32, 256, sm_count() * ctas);
        int nv = (int)((p / W + 31) / 32);
        if (nv == 1)
            launch_cross_runs<F, 1>(n_cat, g, st, X, n, (int)p, d, rows, n_rows, (int)chunk, prm);
        else
            launch_cross_runs<F, 2>(n_cat, g, st, X, n, (int)p, d, rows, n_rows, (int)chunk, prm);
        TM_LAUNCHED();
    } else if (n_rows > 0) {
        int g = grid_for(n_rows * 32, 256, sm_count() * 8);
        int nv = (int)((p / W + 31) / 32);
        if (nv == 1)
            k_dense_cross_fused<F, 1><<<g, 256, 0, st>>>(X, n, (int)p, d, rows, n_rows, prm);
        else
            k_dense_cross_fused<F, 2><<<g, 256, 0, st>>>(X, n, (int)p, d, rows, n_rows, prm);
        TM_LAUNCHED();
    }
    return cross_finish<F>(p, n_cat, K, out_cat, prm, st);
}

int g_cross_runs_mode = 0;
int g_sm_reserve = 0;

template int dense_cross_fused<float>(const float*, int64_t, int64_t, const float*, const int32_t*,
                                      int64_t, int, const int32_t* const*, const int64_t*,
                                      const int32_t*, float* const*, const float*, const int32_t*,
                                      const int32_t*, int64_t, float*, int, cudaStream_t);
template int dense_cross_fused<double>(const double*, int64_t, int64_t, const double*,
                                       const int32_t*, int64_t, int, const int32_t* const*,
                                       const int64_t*, const int32_t*, double* const*,
                                       const double*, const int32_t*, const int32_t*, int64_t,
                                       double*, int, cudaStream_t);

}  // namespace tmb

extern "C" {

void tm_set_cross_runs_mode(int mode) { tmb::g_cross_runs_mode = mode; }
void tm_set_sm_reserve(int sms) { tmb::g_sm_reserve = sms < 0 ? 0 : (sms > 64 ? 64 : sms); }

int tm_csc_dense_gather_sandwich_f32(const float* bdata, const int32_t* brow, const int32_t* bptr,
                                     int64_t p_sparse, int64_t n_blocks, const float* B, int64_t q,
                                     const float* d, float* out, tm_stream_t stream) {
    return tmb::csc_dense_gather<float>(B, q, d, bdata, brow, bptr, p_sparse, n_blocks, out,
                                        tmb::as_stream(stream));
}
int tm_csc_dense_gather_sandwich_f64(const double* bdata, const int32_t* brow,
                                     const int32_t* bptr, int64_t p_sparse, int64_t n_blocks,
                                     const double* B, int64_t q, const double* d, double* out,
                                     tm_stream_t stream) {
    return tmb::csc_dense_gather<double>(B, q, d, bdata, brow, bptr, p_sparse, n_blocks, out,
                                         tmb::as_stream(stream));
}

int tm_dense_cross_sandwich_f32(const float* X, int64_t n, int64_t p, const float* d,
                                const int32_t* rows, int64_t n_rows, int n_cat,
                                const int32_t* const* codes, const int64_t* K,
                                const int32_t* drop_first, float* const* out_cat,
                                const float* csr_data, const int32_t* csr_indices,
                                const int32_t* csr_indptr, int64_t p_sparse, float* out_sparse,
                                tm_stream_t stream) {
    return tmb::dense_cross_fused<float>(X, n, p, d, rows, n_rows, n_cat, codes, K, drop_first,
                                         out_cat, csr_data, csr_indices, csr_indptr, p_sparse,
                                         out_sparse, 1, tmb::as_stream(stream));
}
int tm_dense_cross_sandwich_f64(const double* X, int64_t n, int64_t p, const double* d,
                                const int32_t* rows, int64_t n_rows, int n_cat,
                                const int32_t* const* codes, const int64_t* K,
                                const int32_t* drop_first, double* const* out_cat,
                                const double* csr_data, const int32_t* csr_indices,
                                const int32_t* csr_indptr, int64_t p_sparse, double* out_sparse,
                                tm_stream_t stream) {
    return tmb::dense_cross_fused<double>(X, n, p, d, rows, n_rows, n_cat, codes, K, drop_first,
                                          out_cat, csr_data, csr_indices, csr_indptr, p_sparse,
                                          out_sparse, 1, tmb::as_stream(stream));
}

}  // extern "C"
