// L2 RED.ADD.F32 throughput microbenchmark: every warp adds 128 consecutive floats (one table
// row) per step to a pseudo-random row of a (rows x 128) f32 table; no other memory traffic.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int VEC>
__global__ void k_red(float* table, int rows, long long steps_per_warp, unsigned seed) {
    int lane = threadIdx.x & 31;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned s = seed ^ (unsigned)(warp * 2654435761u);
    for (long long i = 0; i < steps_per_warp; ++i) {
        s = s * 1664525u + 1013904223u;
        int r = (int)((s >> 8) % (unsigned)rows);
        float* row = table + (size_t)r * 128;
        if (VEC == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(row + lane + 32 * j, 1.0f);
        } else {
            float* p = row + lane * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(1.f),
                         "f"(1.f), "f"(1.f)
                         : "memory");
        }
    }
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int rows_list[] = {10, 50, 200, 1280, 1600, 3000, 6260, 100000};
    for (int rows : rows_list) {
        float* table;
        cudaMalloc(&table, (size_t)rows * 128 * 4);
        cudaMemset(table, 0, (size_t)rows * 128 * 4);
        for (int vec = 1; vec <= 4; vec += 3) {
            int blocks = sms * 8, threads = 256;
            long long warps = (long long)blocks * threads / 32;
            long long steps = rows < 100 ? 2000 : 20000;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (vec == 1) k_red<1><<<blocks, threads>>>(table, rows, steps, 123u + rep);
                else k_red<4><<<blocks, threads>>>(table, rows, steps, 123u + rep);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double bytes = (double)warps * steps * 512.0;
            printf("rows=%6d vec=%d  %.3f ms  payload %.2f TB/s  (%.3e row-adds/s) err=%s\n", rows, vec,
                   ms, bytes / (ms * 1e-3) / 1e12, (double)warps * steps / (ms * 1e-3),
                   cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(table);
    }
    return 0;
}
