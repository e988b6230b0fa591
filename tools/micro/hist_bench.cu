// C3 exploration (n = 1e7 rows, K = 2000 levels, f32): where do the 44 us of k_cat_hist2 go?
// Variants of the weighted histogram out[codes[k]] += d[k], each timed alone with CUDA events
// after an L2 flush (a 256 MB fill), 20 repetitions, median.  80 MB of input = 12.2 us at the
// measured 6.55 TB/s.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int T = 1024;

__device__ __forceinline__ bool run_reduce(int key, float& val, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int prev = __shfl_up_sync(FULL, key, 1);
    const bool head = lane == 0 || prev != key;
    const unsigned heads = __ballot_sync(FULL, head);
    if (heads != FULL) {
        const unsigned above = heads & ~((2u << lane) - 1u);
        const int end = above ? __ffs(above) - 2 : 31;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const float o = __shfl_down_sync(FULL, val, off);
            if (lane + off <= end) val += o;
        }
    }
    return head;
}

// MODE 0: stream only (no table)      1: production form (run_reduce + smem atomics + RED flush)
//      2: no run_reduce               3: no flush REDs (plain stores of the CTA's partial table)
//      4: no run_reduce, no flush REDs
template <int MODE, int UNR, int THREADS, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
k_hist(const int* __restrict__ codes, const float* __restrict__ w, long long n, int K, int copies,
       float* __restrict__ out, float* __restrict__ partial) {
    extern __shared__ float table[];
    for (int i = threadIdx.x; i < K * copies; i += THREADS) table[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float* tab = table + (threadIdx.x % copies) * K;
    const long long gw = ((long long)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * THREADS) >> 5;
    float sink = 0.f;
    for (long long base = gw * (32 * UNR); base < n; base += nw * (32 * UNR)) {
        int c[UNR];
        float v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long long t = base + u * 32 + lane;
            c[u] = -1;
            v[u] = 0.f;
            if (t < n) {
                c[u] = codes[t];
                v[u] = w[t];
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (MODE == 0) {
                sink += v[u] + (float)c[u];
                continue;
            }
            int key = c[u];
            bool head = true;
            if (MODE == 1 || MODE == 3) head = run_reduce(key, v[u], lane);
            if (head && key >= 0) atomicAdd(&tab[key], v[u]);
        }
    }
    if (MODE == 0) {
        if (sink == 123.456f) out[0] = sink;
        return;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += THREADS) {
        float s = 0.f;
        for (int r = 0; r < copies; ++r) s += table[r * K + i];
        if (MODE == 3 || MODE == 4)
            partial[(long long)blockIdx.x * K + i] = s;
        else if (s != 0.f)
            atomicAdd(&out[i], s);
    }
}

// vector loads: each thread takes 4 consecutive rows per 16-byte load, V loads in flight
template <int V>
__global__ void __launch_bounds__(T)
k_hist_vec(const int4* __restrict__ codes, const float4* __restrict__ w, long long n4, int K,
           int copies, float* __restrict__ out) {
    extern __shared__ float table[];
    for (int i = threadIdx.x; i < K * copies; i += T) table[i] = 0.f;
    __syncthreads();
    float* tab = table + (threadIdx.x % copies) * K;
    const long long stride = (long long)gridDim.x * T;
    for (long long t = (long long)blockIdx.x * T + threadIdx.x; t < n4; t += stride * V) {
        int4 c[V];
        float4 v[V];
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const long long tt = t + u * stride;
            if (tt < n4) {
                c[u] = codes[tt];
                v[u] = w[tt];
            } else {
                c[u] = make_int4(-1, -1, -1, -1);
            }
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
            if (c[u].x >= 0) atomicAdd(&tab[c[u].x], v[u].x);
            if (c[u].y >= 0) atomicAdd(&tab[c[u].y], v[u].y);
            if (c[u].z >= 0) atomicAdd(&tab[c[u].z], v[u].z);
            if (c[u].w >= 0) atomicAdd(&tab[c[u].w], v[u].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += T) {
        float s = 0.f;
        for (int r = 0; r < copies; ++r) s += table[r * K + i];
        if (s != 0.f) atomicAdd(&out[i], s);
    }
}

int main() {
    const long long n = 10000000;
    const int K = 2000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<int> hc(n);
    std::vector<float> hw(n);
    unsigned s = 12345u;
    for (long long i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        hc[i] = (int)((s >> 8) % K);
        hw[i] = (float)((s >> 4) & 1023) / 1024.f;
    }
    int* codes;
    float *w, *out, *partial, *flush;
    cudaMalloc(&codes, n * 4);
    cudaMalloc(&w, n * 4);
    cudaMalloc(&out, K * 4);
    cudaMalloc(&partial, (size_t)sms * 4 * K * 4);
    const size_t flush_bytes = 256u << 20;
    cudaMalloc(&flush, flush_bytes);
    cudaMemcpy(codes, hc.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(w, hw.data(), n * 4, cudaMemcpyHostToDevice);
    std::vector<double> ref(K, 0.0);
    for (long long i = 0; i < n; ++i) ref[hc[i]] += hw[i];
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto bench = [&](const char* name, auto launch, bool check) {
        std::vector<float> ts;
        for (int rep = 0; rep < 22; ++rep) {
            cudaMemsetAsync(flush, rep, flush_bytes);
            cudaMemsetAsync(out, 0, K * 4);
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2) ts.push_back(ms);
        }
        std::sort(ts.begin(), ts.end());
        double err = -1;
        if (check) {
            std::vector<float> h(K);
            cudaMemcpy(h.data(), out, K * 4, cudaMemcpyDeviceToHost);
            err = 0;
            for (int i = 0; i < K; ++i) err = std::max(err, std::abs(h[i] - ref[i]) / std::abs(ref[i]));
        }
        printf("%-58s median %7.2f us  min %7.2f us  (%.0f GB/s)  relerr %.1e  %s\n", name,
               ts[ts.size() / 2] * 1e3, ts[0] * 1e3, 80.008e6 / (ts[ts.size() / 2] * 1e-3) / 1e9, err,
               cudaGetErrorString(cudaGetLastError()));
    };
    const int copies = 6;
    const size_t sm = (size_t)K * copies * 4;
    bench("stream only, 148 x 1024, UNR 8", [&] { k_hist<0, 8, 1024><<<sms, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("stream only, 296 x 1024 (2 CTA/SM), UNR 4", [&] { k_hist<0, 4, 1024, 2><<<sms * 2, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("stream only, 148 x 1024, UNR 16", [&] { k_hist<0, 16, 1024><<<sms, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("production form (run_reduce, 6 copies, RED flush)", [&] { k_hist<1, 8, 1024><<<sms, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, true);
    bench("  without run_reduce", [&] { k_hist<2, 8, 1024><<<sms, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, true);
    bench("  without flush REDs (partial tables stored)", [&] { k_hist<3, 8, 1024><<<sms, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("  without both", [&] { k_hist<4, 8, 1024><<<sms, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("  without both, 1 copy", [&] { k_hist<4, 8, 1024><<<sms, 1024, (size_t)K * 4>>>(codes, w, n, K, 1, out, partial); }, false);
    bench("  without both, UNR 4, 2 CTA/SM x 1024", [&] { k_hist<4, 4, 1024, 2><<<sms * 2, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("  without both, UNR 16", [&] { k_hist<4, 16, 1024><<<sms, 1024, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("  without both, 512 threads x 4 CTA/SM, UNR 4", [&] { k_hist<4, 4, 512, 4><<<sms * 4, 512, sm>>>(codes, w, n, K, copies, out, partial); }, false);
    bench("vector loads (int4/float4), 1 in flight, RED flush", [&] { k_hist_vec<1><<<sms, 1024, sm>>>((const int4*)codes, (const float4*)w, n / 4, K, copies, out); }, true);
    bench("vector loads (int4/float4), 2 in flight, RED flush", [&] { k_hist_vec<2><<<sms, 1024, sm>>>((const int4*)codes, (const float4*)w, n / 4, K, copies, out); }, true);
    bench("vector loads, 2 in flight, 2 CTA/SM", [&] { k_hist_vec<2><<<sms * 2, 1024, sm>>>((const int4*)codes, (const float4*)w, n / 4, K, copies, out); }, true);
    return 0;
}
