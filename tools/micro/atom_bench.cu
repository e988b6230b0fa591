// Scalar float atomic throughput on B200: shared-memory atomicAdd (ATOMS.CAST.SPIN loops),
// per-thread private shared-memory RMW, and scalar L2 REDs, each thread adding to a
// pseudo-random entry of a table of S floats.  Prints lane-adds per second.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned lcg(unsigned& s) {
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}

// mode 0: shared atomics, one table; 1: shared atomics, table replicated per lane%rep;
// 2: private [S][T] plain RMW
template <int MODE>
__global__ void __launch_bounds__(1024) k_smem(float* out, int S, int rep, int steps) {
    extern __shared__ float tab[];
    const int T = blockDim.x;
    const int total = MODE == 2 ? S * T : S * rep;
    for (int i = threadIdx.x; i < total; i += T) tab[i] = 0.f;
    __syncthreads();
    unsigned s = 12345u ^ (unsigned)((blockIdx.x * T + threadIdx.x) * 2654435761u);
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < steps; ++i) {
        const int k = (int)(lcg(s) % (unsigned)S);
        if (MODE == 2)
            tab[k * T + threadIdx.x] += 1.f;
        else
            atomicAdd(&tab[(lane % rep) * S + k], 1.f);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = tab[0];
}

__global__ void __launch_bounds__(1024) k_l2(float* table, long long S, int steps) {
    unsigned s = 777u ^ (unsigned)((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u);
    for (int i = 0; i < steps; ++i) {
        unsigned a = lcg(s), b = lcg(s);
        long long k = (long long)((((unsigned long long)a << 24) ^ b) % (unsigned long long)S);
        atomicAdd(table + k, 1.f);
    }
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    float* out;
    cudaMalloc(&out, 1 << 20);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int steps = 4000;
    auto report = [&](const char* what, int S, int rep, int threads, float ms) {
        double adds = (double)sms * threads * steps;
        printf("%-28s S=%8d rep=%2d thr=%4d  %.3f ms  %.3e adds/s  (%.2f lane-adds/clk/SM at 1.9 GHz) %s\n",
               what, S, rep, threads, ms, adds / (ms * 1e-3), adds / (ms * 1e-3) / sms / 1.9e9,
               cudaGetErrorString(cudaGetLastError()));
    };
    const int S_list[] = {10, 50, 200, 500, 2000, 10000, 40000};
    for (int S : S_list) {
        for (int rep : {1, 8, 32}) {
            if ((size_t)S * rep * 4 > 190 * 1024) continue;
            for (int it = 0; it < 2; ++it) {
                cudaEventRecord(e0);
                k_smem<1><<<sms, 1024, (size_t)S * rep * 4>>>(out, S, rep, steps);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            report("smem atomicAdd", S, rep, 1024, ms);
        }
        if ((size_t)S * 1024 * 4 <= 190 * 1024) {
            for (int it = 0; it < 2; ++it) {
                cudaEventRecord(e0);
                k_smem<2><<<sms, 1024, (size_t)S * 1024 * 4>>>(out, S, 1, steps);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            report("smem private RMW", S, 1024, 1024, ms);
        }
    }
    const long long L_list[] = {500, 2000, 10000, 20000, 50000, 200000, 2000000, 20000000, 200000000};
    for (long long S : L_list) {
        float* table;
        cudaMalloc(&table, (size_t)S * 4);
        cudaMemset(table, 0, (size_t)S * 4);
        for (int blocks_per_sm : {1, 2}) {
            for (int it = 0; it < 2; ++it) {
                cudaEventRecord(e0);
                k_l2<<<sms * blocks_per_sm, 1024>>>(table, S, steps / 4);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double adds = (double)sms * blocks_per_sm * 1024 * (steps / 4);
            printf("L2 RED scalar               S=%10lld ctas/SM=%d  %.3f ms  %.3e adds/s %s\n", S,
                   blocks_per_sm, ms, adds / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(table);
    }
    return 0;
}
