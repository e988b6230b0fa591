// Round-2 question: can the dense x sparse scatter (one 512-byte row-add per non-zero into a
// 3000 x 128 f32 table, today 1.17e10 row-adds/s = 6.0 TB/s of L2 vector REDs) run faster with
// the table spread over the shared memory of a thread-block cluster?  Every warp adds a
// 128-float row (4 floats per lane) to a pseudo-random table row; the owner of a row is the
// CTA `row / rows_per_cta` of the cluster.
//   mode 0: local shared memory only (rows of the own CTA), atomicAdd            (CAS loops)
//   mode 1: cluster-wide, generic pointer from cluster.map_shared_rank + atomicAdd
//   mode 2: cluster-wide, mapa + red.relaxed.cluster.shared::cluster.add.f32
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned lcg(unsigned& s) {
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}

template <int MODE>
__global__ void __launch_bounds__(256)
k_dsmem(float* out, int rows_per_cta, int steps) {
    extern __shared__ __align__(16) float tab[];  // rows_per_cta x 128
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    for (int i = threadIdx.x; i < rows_per_cta * 128; i += blockDim.x) tab[i] = 0.f;
    cluster.sync();
    const int lane = threadIdx.x & 31;
    unsigned s = 99u ^ (unsigned)(((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2654435761u);
    const int total_rows = rows_per_cta * (MODE == 0 ? 1 : csize);
    for (int i = 0; i < steps; ++i) {
        const int r = (int)(lcg(s) % (unsigned)total_rows);
        const int owner = MODE == 0 ? 0 : r / rows_per_cta;
        const int lr = MODE == 0 ? r : r - owner * rows_per_cta;
        float* local = tab + lr * 128 + lane * 4;
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(local + j, 1.f);
        } else if (MODE == 1) {
            float* remote = cluster.map_shared_rank(local, owner);
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(remote + j, 1.f);
        } else {
            unsigned la = (unsigned)__cvta_generic_to_shared(local), ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(owner));
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" ::"r"(ra + 4 * j),
                             "f"(1.f)
                             : "memory");
        }
    }
    cluster.sync();
    if (threadIdx.x == 0) out[blockIdx.x] = tab[0];
}

template <int MODE>
static void run(const char* name, int csize, int rows_per_cta, int sms, float* out) {
    const int steps = 2000;
    const size_t smem = (size_t)rows_per_cta * 512;
    cudaFuncSetAttribute(k_dsmem<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_dsmem<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int grid = sms / csize * csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t err = cudaSuccess;
    float ms = 0;
    for (int it = 0; it < 2; ++it) {
        cudaEventRecord(e0);
        err = cudaLaunchKernelEx(&cfg, k_dsmem<MODE>, out, rows_per_cta, steps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaError_t err2 = cudaGetLastError();
    double row_adds = (double)grid * 8 * steps;
    printf("%-44s cluster %2d, %3d rows/CTA (%3zu KB), grid %3d: %8.3f ms  %.3e row-adds/s = %.2f TB/s of payload  %s %s\n",
           name, csize, rows_per_cta, smem / 1024, grid, ms, row_adds / (ms * 1e-3),
           row_adds * 512 / (ms * 1e-3) / 1e12, cudaGetErrorString(err), cudaGetErrorString(err2));
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, 4096);
    run<0>("local smem atomicAdd", 1, 188, sms, out);
    run<0>("local smem atomicAdd", 1, 375, sms, out);
    run<1>("DSMEM generic atomicAdd", 8, 375, sms, out);
    run<2>("DSMEM red.shared::cluster", 8, 375, sms, out);
    run<1>("DSMEM generic atomicAdd", 16, 188, sms, out);
    run<2>("DSMEM red.shared::cluster", 16, 188, sms, out);
    run<2>("DSMEM red.shared::cluster", 4, 375, sms, out);
    run<2>("DSMEM red.shared::cluster", 2, 375, sms, out);
    printf("reference: L2 vector REDs sustain 1.17e10 row-adds/s = 6.0 TB/s (tools/micro/red_bench.cu)\n");
    return 0;
}
