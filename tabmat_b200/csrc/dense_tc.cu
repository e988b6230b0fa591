// Dense fp32 weighted SYRK  out = X^T diag(d) X  on the 5th-gen tensor cores (sm_100a):
// tcgen05.mma kind::tf32 with accumulators in TMEM, X tiles staged by TMA, d folded in
// between the HBM->smem stage and the MMA (the reference folds d while packing its R
// panel, dense_helpers-tmpl.cpp:224,229).
//
// Data flow per CTA (one persistent CTA per SM, split-K over row tiles of BK=32 rows):
//
//   warp 0  (1 lane)  TMA producer: X[k0:k0+32, :] -> smem stage "R" as column groups of
//                     32 floats (one swizzled box [32 rows][128 B] per group).  For row-major
//                     X this is exactly the MN-major SWIZZLE_128B_BASE32B canonical UMMA
//                     layout (4-row x 128 B atoms stacked along K), so no transpose is needed.
//   warps 2-9         scale warps: B[r][c] = rna_tf32(d[k0+r] * R[r][c]) into a second smem
//                     buffer with the same (swizzled) addresses; R is rounded to tf32 in place
//                     (round-to-nearest; raw fp32 bits would be truncated by the MMA, a biased
//                     error).  fence.proxy.async, then arrive on the stage's "scaled" barrier.
//   warp 1  (1 lane)  MMA issuer: for each 8-row K step and each lower-triangular 128x128
//                     output tile (mt >= nt):  D[mt,nt] += R[:,mt]^T * B[:,nt]
//                     (A = R, B = scaled, both MN-major); tcgen05.commit frees the stage.
//   warps 2-9         epilogue: tcgen05.ld the accumulators and RED.ADD them into `out`
//                     transposed (upper triangle, coalesced across the warp's lanes).
//
// A small second kernel mirrors the upper triangle into the lower one.
// Only the C-order, P <= 256 case is handled here; everything else is served by the
// CUDA-core kernel in dense.cu.
#include <cuda.h>

#include "tm_common.cuh"

namespace tmb {

int g_dense_f32_mode = 0;

namespace tc {

constexpr int BK = 32;                     // rows per pipeline stage
constexpr int GROUP_BYTES = BK * 128;      // one 32-column group of a stage: [BK][128 B]
constexpr int NUM_SCALE_WARPS = 8;
constexpr int NUM_SCALE_THREADS = NUM_SCALE_WARPS * 32;
constexpr int NUM_THREADS = 64 + NUM_SCALE_THREADS;  // producer warp + mma warp + scale warps
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 200 * 1024;

struct Params {
    const float* d;
    float* out;
    long long n;
    int P;            // number of columns (<= 256)
    int groups;       // ceil(P/32): column groups actually loaded by TMA
    int mtiles;       // ceil(P/128): 128-wide output tile rows (1 or 2)
    long long num_row_tiles;
    int stages;
    int tmem_cols;
    float* dbg;       // debug dump (tools/tc_debug.py); nullptr in production
    int variant;      // bring-up switches (0 in production)
};

// ---- PTX wrappers ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                 uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell) | [61,64) layout type (1 = SWIZZLE_128B_BASE32B)
// MN-major 32-bit operands only exist in the 128B-swizzle-with-32B-atoms layout
// (Swizzle<2,5,2>: 32-byte chunk index ^= row & 3; K atom = 4 rows of 128 B = 512 B), which is
// what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  One K=8 instruction therefore
// spans two K atoms: SBO = 512 B; the next 32-column group is LBO = GROUP_BYTES away.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B: the only MN-major layout tf32 supports
    return d;
}

// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32,
// both operands MN-major, M=128, N=128.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TM_TMEM_LD_32x32B_X32(taddr, v)                                                            \
    asm volatile(                                                                                  \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                  \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                  \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"  \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),      \
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),  \
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),            \
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),            \
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])             \
        : "r"(taddr))

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_dense_syrk_tc(const __grid_constant__ CUtensorMap tmap, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int S = prm.stages;
    const int G = prm.mtiles * 4;  // column groups allocated per stage (multiple of 4)
    const uint32_t half_bytes = (uint32_t)G * GROUP_BYTES;
    const uint32_t stage_bytes = 2 * half_bytes;

    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* scaled = bars + MAX_STAGES;
    uint64_t* empty = bars + 2 * MAX_STAGES;
    uint64_t* done = bars + 3 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // Column groups that TMA never writes (P not a multiple of 128) must read as zeros.
    if (prm.groups < G) {
        for (int s = 0; s < S; ++s) {
            for (int h = 0; h < 2; ++h) {
                uint4* z = reinterpret_cast<uint4*>(base + (size_t)s * stage_bytes + h * half_bytes +
                                                    (size_t)prm.groups * GROUP_BYTES);
                int cnt = (G - prm.groups) * GROUP_BYTES / 16;
                for (int i = threadIdx.x; i < cnt; i += NUM_THREADS) z[i] = make_uint4(0, 0, 0, 0);
            }
        }
        fence_proxy_async();
    }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&scaled[s], NUM_SCALE_THREADS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmap) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"((uint32_t)prm.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // row tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    long long my_count = 0;
    if ((long long)blockIdx.x < prm.num_row_tiles)
        my_count = (prm.num_row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (long long it = 0; it < my_count; ++it) {
                int s = (int)(it % S);
                uint32_t ph = (uint32_t)((it / S) & 1);
                mbar_wait(&empty[s], ph ^ 1);
                long long k0 = ((long long)blockIdx.x + it * gridDim.x) * BK;
                uint8_t* R = base + (size_t)s * stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)prm.groups * GROUP_BYTES);
                for (int g = 0; g < prm.groups; ++g)
                    tma_load_2d(R + (size_t)g * GROUP_BYTES, &tmap, &full[s], g * 32, (int)k0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(128, 128, 1, 1);
        for (long long it = 0; it < my_count; ++it) {
            int s = (int)(it % S);
            uint32_t ph = (uint32_t)((it / S) & 1);
            mbar_wait(&scaled[s], ph);
            tcgen05_fence_after();
            if (lane == 0) {
                uint32_t Ra = smem_u32(base + (size_t)s * stage_bytes);
                uint32_t Ba = Ra + half_bytes;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    int tile = 0;
                    for (int mt = 0; mt < prm.mtiles; ++mt) {
                        const uint32_t lbo = (prm.variant & 1) ? 512u : (uint32_t)GROUP_BYTES;
                        const uint32_t sbo = (prm.variant & 1) ? (uint32_t)GROUP_BYTES : 512u;
                        uint64_t da = make_desc(Ra + (uint32_t)(mt * 4) * GROUP_BYTES + ks * 1024,
                                                lbo, sbo);
                        for (int nt = 0; nt <= mt; ++nt, ++tile) {
                            uint64_t db = make_desc(
                                Ba + (uint32_t)(nt * 4) * GROUP_BYTES + ks * 1024, lbo, sbo);
                            tcgen05_mma_tf32(tmem_base + (uint32_t)tile * 128, da, db, idesc, acc);
                        }
                    }
                }
                tcgen05_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (lane == 0) tcgen05_commit(done);
        __syncwarp();
    } else {
        // ===== scale warps, then epilogue =====
        const int t = (int)threadIdx.x - 64;  // 0..255
        const int r = t >> 3;                 // row of this thread's 16-byte chunk inside a stage
        float d_next = 0.f;
        if (my_count > 0) {
            long long k = (long long)blockIdx.x * BK + r;
            d_next = (k < prm.n) ? prm.d[k] : 0.f;
        }
        for (long long it = 0; it < my_count; ++it) {
            int s = (int)(it % S);
            uint32_t ph = (uint32_t)((it / S) & 1);
            float dk = d_next;
            if (it + 1 < my_count) {
                long long k = ((long long)blockIdx.x + (it + 1) * gridDim.x) * BK + r;
                d_next = (k < prm.n) ? prm.d[k] : 0.f;
            }
            mbar_wait(&full[s], ph);
            uint8_t* R = base + (size_t)s * stage_bytes + (size_t)t * 16;
            uint8_t* B = R + half_bytes;
            for (int g = 0; g < prm.groups; ++g) {
                float4 x = *reinterpret_cast<const float4*>(R + (size_t)g * GROUP_BYTES);
                uint4 a, b;
                a.x = to_tf32(x.x);
                a.y = to_tf32(x.y);
                a.z = to_tf32(x.z);
                a.w = to_tf32(x.w);
                b.x = to_tf32(dk * x.x);
                b.y = to_tf32(dk * x.y);
                b.z = to_tf32(dk * x.z);
                b.w = to_tf32(dk * x.w);
                *reinterpret_cast<uint4*>(R + (size_t)g * GROUP_BYTES) = a;
                *reinterpret_cast<uint4*>(B + (size_t)g * GROUP_BYTES) = b;
            }
            fence_proxy_async();
            if (prm.dbg && blockIdx.x == 0 && it == 0) {
                // raw stage-0 smem (R half then B half) as floats
                const float* sm = reinterpret_cast<const float*>(base);
                for (uint32_t i = t; i < stage_bytes / 4; i += NUM_SCALE_THREADS) prm.dbg[i] = sm[i];
            }
            mbar_arrive(&scaled[s]);
        }

        // epilogue: TMEM -> registers -> RED into the upper triangle of `out` (transposed)
        mbar_wait(done, 0);
        tcgen05_fence_after();
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int chalf = (warp - 2) >> 2;    // which 64-column half of a tile this warp drains
        const int P = prm.P;
        if (my_count > 0) {
            int tile = 0;
            for (int mt = 0; mt < prm.mtiles; ++mt) {
                for (int nt = 0; nt <= mt; ++nt, ++tile) {
                    const int C = mt * 128 + q * 32 + lane;  // output column (= X column of A)
#pragma unroll 1
                    for (int cc = 0; cc < 2; ++cc) {
                        const int n0 = chalf * 64 + cc * 32;
                        uint32_t v[32];
                        uint32_t taddr =
                            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * 128 + n0);
                        TM_TMEM_LD_32x32B_X32(taddr, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (prm.dbg && blockIdx.x == 0) {
                            float* T = prm.dbg + 65536 + (size_t)tile * 128 * 128;
                            for (int j = 0; j < 32; ++j)
                                T[(q * 32 + lane) * 128 + n0 + j] = __uint_as_float(v[j]);
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int Rr = nt * 128 + n0 + j;  // output row (= X column of B)
                            if (Rr < P && C < P && C >= Rr)
                                atomicAdd(&prm.out[(size_t)Rr * P + C], __uint_as_float(v[j]));
                        }
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)prm.tmem_cols)
                     : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

static int device_cc_major() {
    static int cc = -1;
    if (cc < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cc = major;
    }
    return cc;
}

}  // namespace tc

bool dense_tc_eligible(int64_t n, int64_t p, int c_order, const void* X) {
    if (!c_order) return false;
    if (p < 8 || p > 256 || (p % 4) != 0) return false;     // TMA: row pitch multiple of 16 B
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0) return false;
    if (n < 1 || n > 0x7fffffffLL) return false;
    if (tc::device_cc_major() != 10) return false;
    return tc::get_encode() != nullptr;
}

float* g_tc_dbg = nullptr;
int g_tc_variant = 0;

int dense_sandwich_tc_f32(const float* X, int64_t n, int64_t p, int c_order, const float* d,
                          float* out, cudaStream_t st) {
    using namespace tc;
    (void)c_order;
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail("cuTensorMapEncodeTiled not available");

    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)p, (cuuint64_t)n};
    cuuint64_t gstride[1] = {(cuuint64_t)p * sizeof(float)};
    cuuint32_t box[2] = {32, (cuuint32_t)BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim,
                      gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed");

    Params prm;
    prm.d = d;
    prm.out = out;
    prm.n = n;
    prm.P = (int)p;
    prm.groups = (int)((p + 31) / 32);
    prm.mtiles = (int)((p + 127) / 128);
    prm.num_row_tiles = (n + BK - 1) / BK;
    prm.dbg = g_tc_dbg;
    prm.variant = g_tc_variant;
    int ntiles = prm.mtiles * (prm.mtiles + 1) / 2;
    prm.tmem_cols = ntiles == 1 ? 128 : 512;
    int stage_bytes = 2 * prm.mtiles * 4 * GROUP_BYTES;
    int stages = (SMEM_BUDGET - 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return fail("dense_tc: not enough shared memory for 2 stages");
    prm.stages = stages;
    size_t smem = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 512 /*barriers*/;

    static bool attr_set = false;
    if (!attr_set) {
        TM_CUDA(cudaFuncSetAttribute(k_dense_syrk_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
        attr_set = true;
    }
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(p * p), st));
    long long grid = prm.num_row_tiles < sm_count() ? prm.num_row_tiles : sm_count();
    k_dense_syrk_tc<<<(unsigned)grid, NUM_THREADS, smem, st>>>(tmap, prm);
    TM_LAUNCHED();
    return symmetrize_from_upper<float>(out, p, st);
}

}  // namespace tmb

extern "C" {

int tm_has_tcgen05(void) {
    return tmb::tc::device_cc_major() == 10 && tmb::tc::get_encode() != nullptr ? 1 : 0;
}
void tm_set_dense_f32_mode(int mode) { tmb::g_dense_f32_mode = mode; }
/* test hook (not part of the public header): device buffer of >= 65536 + 3*128*128 floats */
void tm_debug_set_tc_buffer(float* buf) { tmb::g_tc_dbg = buf; }
void tm_debug_set_tc_variant(int v) { tmb::g_tc_variant = v; }

}  // extern "C"
