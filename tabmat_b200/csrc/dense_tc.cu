// Dense fp32 weighted SYRK  out = X^T diag(d) X  on the 5th-gen tensor cores (sm_100a):
// tcgen05.mma kind::tf32 with accumulators in TMEM, X tiles staged by TMA, d folded in
// between the HBM->smem stage and the MMA (the reference folds d while packing its R
// panel, dense_helpers-tmpl.cpp:224,229).  Categorical blocks with few levels ride along
// as one-hot MMAs against the same d-scaled tile (the reference makes one
// sandwich_cat_dense call per block, split.pyx:32-80).
//
// Data flow per CTA (one persistent CTA per SM, split-K over row tiles of BK=32 rows):
//
//   warp 0  (1 lane)  TMA producer: one box X[k0:k0+32, 0:P] per stage -> ring "R"
//                     (row-major, no swizzle; only the scale warps read it).
//   warps 2-9         scale warps: transpose + round + scale each stage into the K-major
//                     SWIZZLE_128B UMMA layout (one 128-byte row of 32 k-values per X column):
//                       A'[c][k] = rna_tf32(X[k0+k, c])         B'[c][k] = rna_tf32(d[k0+k] * X[k0+k, c])
//                     and set the one-hot operand  O[slot][k] = 1  (slot = off_i + code_i[k0+k]);
//                     reads are conflict-free LDS.32 (lanes along c), writes are STS.128 of
//                     4 consecutive k.  (MN-major tf32 operands straight from the TMA tile work
//                     too - SWIZZLE_128B_BASE32B - but issue 4-6x slower than the 64-cycle MMA
//                     floor on B200, measured; K-major runs at the floor.)
//   warp 1  (1 lane)  MMA issuer, per 8-row K step:
//                       D[mt,nt]   += A'[mt] * B'[nt]^T   for the lower-triangular 128x128 tiles
//                       D'[c,slot] += B'[0]  * O^T        one-hot blocks (N up to 256 per MMA)
//                     tcgen05.commit releases the R stage and the A'/B'/O slot.
//   warps 2-9         epilogue: tcgen05.ld the accumulators and RED.ADD them into `out`
//                     (transposed: lanes = output columns, coalesced) / into the one-hot result.
//
// A small second kernel mirrors the upper triangle into the lower one.
// Only the C-order, P <= 256 case is handled here; everything else is served by the
// CUDA-core kernel in dense.cu.
#include <cuda.h>

#include <cstdlib>

#include "tm_common.cuh"

namespace tmb {

int g_dense_f32_mode = 0;

namespace tc {

constexpr int BK = 32;                     // rows per pipeline stage (= one 128 B K-major row)
constexpr int TILE_BYTES = 128 * 128;      // one K-major operand tile: 128 X-columns x 32 k x 4 B
constexpr int GROUP_BYTES = 32 * 128;      // 32 one-hot slots x 128 B
// producer warp + mma warp + NSW scale warps (8, or 16: template parameter of the kernel) + SCW
// scatter warps
constexpr int tc_threads(int nsw, int scw) { return 64 + 32 * nsw + 32 * scw; }
constexpr int MAX_STAGES = 16;
constexpr int MAX_SB = 4;                  // max depth of the operand (S / one-hot / T) ring
constexpr int SMEM_BUDGET = 220 * 1024;

struct Params {
    const float* d;
    float* out;
    long long n;
    int P;            // number of columns (<= 256)
    int groups;       // ceil(P/32)
    int mtiles;       // ceil(P/128): 128-wide output tile rows (1 or 2)
    long long num_row_tiles;   // row tiles of THIS launch
    long long tile0;           // first row tile of this launch (long inputs are cut into segments)
    int stagesR;      // depth of the TMA ring (raw X tiles)
    int sb;           // depth of the operand ring (2..MAX_SB): S + one-hot in smem, T in TMEM
    int dual_acc;     // mtiles == 1 without one-hot blocks: two SYRK accumulators (even/odd k steps)
    int r_bytes;      // bytes of one raw stage: X tile | d (128 B) | one-hot codes (8 x 128 B)
    int aux_off;      // offset of d inside a stage (BK * P * 4 rounded up to 128)
    int tmem_cols;
    // one-hot extension (P <= 128 only): categorical blocks with few levels ride along as
    // extra MMAs  D'[dense col, slot] += (d*X)^T OneHot,  slot = oh_off[c] + code - oh_df[c]
    int oh_groups;    // ceil(oh_slots / 32); 0 = no one-hot blocks
    int oh_slots;
    int oh_ncat;
    const int32_t* oh_codes[8];
    int oh_off[8];
    int oh_K[8];
    int oh_df[8];
    float* oh_out;    // [oh_slots][P], accumulated with RED
    float* dbg;       // debug dump (tools/tc_debug.py); nullptr in production
    int variant;      // bring-up switches (0 in production)
    // fused scatter work (SCW > 0 only; P <= 128): extra warps read the raw stage and issue the
    // vector REDs of dense x sparse and dense x many-level categoricals (split_fused.cu does the
    // same from a second pass over X)
    int f_order;      // X is column-major: the TMA box lands K-major (SWIZZLE_128B) in the stage
    // fused IRLS pass: vec_out[c] += sum_k v[k] * X[k, c] in full fp32 on the CUDA cores of the
    // scale warps (v is staged by TMA next to d); has_v = 0: plain sandwich
    int has_v;
    float* vec_out;
    // 1: one TF32 MMA pass per stage.  3: "3xTF32" - every fp32 operand is split into
    // hi = tf32(a) and lo = tf32(a - hi) and the products hi*hi + hi*lo + lo*hi are accumulated
    // (three operand slots per raw stage), which restores fp32-level accuracy at a third of the
    // tensor throughput (tm_set_dense_f32_mode(3))
    int nsub;
    int round_mode;   // fp32 -> tf32 of the MMA operands: see to_tf32<RM>
    // column panels.  Legacy (panel == 0): the kernel's P columns are columns 0..P-1 of X and one
    // TMA box brings a whole row tile.  panel == 1 (p > 256): the kernel works on TWO 128-column
    // panels of a wider X - panel A = columns a0..a0+wa-1 (local columns 0..127), panel B =
    // b0..b0+wb-1 (local columns 128..255), one TMA box each - and writes the output tiles
    // selected by tile_mask (bit 0: A x A, bit 1: B x A, bit 2: B x B) at their global position
    // in the ldo x ldo result.  The host loops over panel pairs.
    int panel, a0, b0, wa, wb, ldo, tile_mask;
    int sc_ncat;                       // <= TC_SCATTER_MAX_CATS
    int sc_staged;    // 1: the scatter blocks' code vectors are TMA-staged next to the one-hot
                      // ones (slots oh_ncat .. oh_ncat + sc_ncat - 1 of the stage): the scatter
                      // warps then touch no global memory except for their REDs
    const int32_t* sc_codes[TC_SCATTER_MAX_CATS];
    float* sc_tab[TC_SCATTER_MAX_CATS];
    int sc_K[TC_SCATTER_MAX_CATS];
    int sc_copies[TC_SCATTER_MAX_CATS];
    int sc_df[TC_SCATTER_MAX_CATS];
    const float* csr_data;
    const int32_t* csr_indices;
    const int32_t* csr_indptr;
    float* out_sparse;                 // [p_sparse][P] or nullptr
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
// Arrive WITHOUT release semantics.  The default (.release) makes the warp wait until its prior
// memory operations are performed - including global REDs still in flight (~2000 cycles each
// under load), which is what made the scatter warps of the fused form 4-8x slower than the MMA
// pipeline (measured: 27.7 ms with 4 warps although they issued < 2 REDs per tile).  The only
// ordering the raw-stage ring needs from a consumer is "my shared-memory READS of the stage are
// done", which the callers guarantee by arriving after the loaded values have been used.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// relaxed arrive that the hardware may not issue before `dep` exists: `dep` is computed from the
// values loaded from the stage, so the arrive cannot overtake those shared-memory reads
__device__ __forceinline__ void mbar_arrive_relaxed_after(uint64_t* bar, uint32_t dep) {
    asm volatile(
        "{\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %1, 0;\n\t"
        "add.u32 t, t, %0;\n\t"
        "mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [t];\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(dep)
        : "memory");
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
// The same with a suspend-time hint: the waiting thread sleeps in hardware until the phase
// completes (or the hint expires) instead of re-issuing TRYWAIT + BRA every few cycles.  ncu on
// the fused form showed 4e8 executions of one such spin loop against 1e6 for the useful
// instructions: the spinning scale warps were taking the issue slots of their scheduler away
// from the scatter warp that shares it.  Used by the scale / scatter warps (many waiters); the
// single MMA issuer and the producer keep the tight form.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP_S:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE_S;\n\t"
        "bra WAIT_LOOP_S;\n\t"
        "DONE_S:\n\t"
        "}" ::"r"(addr),
        "r"(parity), "r"(20000u)
        : "memory");
}
template <bool SLEEP>
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity) {
    if (SLEEP)
        mbar_wait_sleep(bar, parity);
    else
        mbar_wait(bar, parity);
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
__device__ __forceinline__ void tma_load_1d(uint32_t dst_sa, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0) {
    asm volatile(
        "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3}], [%2];" ::"r"(dst_sa),
        "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0)
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
// A operand from TMEM (lanes = M rows, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void tcgen05_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a,
                                                    uint64_t desc_b, uint32_t idesc,
                                                    uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// each lane writes 4 consecutive TMEM columns of its own TMEM lane (warp's lane quarter)
__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, uint32_t v0, uint32_t v1, uint32_t v2,
                                           uint32_t v3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr),
                 "r"(v0), "r"(v1), "r"(v2), "r"(v3)
                 : "memory");
}
// explicit shared-space accesses on 32-bit shared addresses (pointers carved out of the dynamic
// shared buffer otherwise degrade to generic LD.E / ST.E)
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32x4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_s32(uint32_t a, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// one elected lane of a converged warp (the compiler then emits the tcgen05 instructions
// straight-line instead of a per-active-lane serialisation loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// fp32 -> tf32, round to nearest, ties away from zero.  ptxas expands `cvt.rna.tf32.f32` into FOUR
// instructions on sm_100a (FSETP |x| >= inf, VIADD 0x1000, SEL, LOP3 & 0xffffe000) and the scale
// warps - which pace the kernel - do 32 conversions per thread and tile.  RM selects:
//   0  the cvt instruction;
//   1  (bits + 0x1000) & 0xffffe000: the same value for every finite x and for +-inf (the carry
//      stops in bit 12, which the mask drops); a NaN stays a NaN unless its payload is all ones;
//   2  bits + 0x1000 only: the tensor core ignores the low 13 bits of a tf32 operand, so the
//      mask is redundant for an MMA operand (NOT for the residual of the 3xTF32 scheme).
template <int RM = 0>
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    if (RM == 0) {
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    } else if (RM == 1) {
        r = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    } else {
        r = __float_as_uint(x) + 0x1000u;
    }
    return r;
}

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// operand of sub-pass `sub` of the 3xTF32 scheme: want_lo selects tf32(a - tf32(a))
template <int RM>
__device__ __forceinline__ uint32_t tf32_part(float a, bool want_lo) {
    if (!want_lo) return to_tf32<RM>(a);
    const uint32_t hi = to_tf32 < RM == 2 ? 1 : RM > (a);
    return to_tf32<RM>(a - __uint_as_float(hi));
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell) | [61,64) layout type (2 = SWIZZLE_128B)
// K-major SWIZZLE_128B: one operand row (an X column / a one-hot slot) = 128 B = 32 k-values,
// 16-byte chunk index ^= row & 7; 8-row groups are SBO = 1024 B apart; LBO is unused.  A K=8
// instruction reads 32 B of every row: the k-step advances the start address by 32 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// byte offset of element (row, k) inside a K-major SWIZZLE_128B tile; k4 = k / 4 (16 B chunk)
__device__ __forceinline__ uint32_t kmajor_chunk_off(uint32_t row, uint32_t k4) {
    return row * 128u + ((k4 ^ (row & 7u)) << 4);
}

// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
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

// Transpose + round + scale 4 row-chunks (k4 = kb, kb+ks, kb+2ks, kb+3ks; 4 rows each) of X
// column c of a raw stage: the rounded values go to the K-major shared-memory tile S (N side of
// the MMAs), the d-scaled values to this lane's TMEM lane of the T operand (M side).  All 16
// loads are issued before the math.  Lanes with c >= P write nothing to S (those rows stay
// zero) and zeros to T (tcgen05.st is warp-collective).
// PITCH: the row pitch of the raw stage in bytes as a compile-time constant (512 = 128 columns,
// 1024 = 256 columns: the 16 row offsets of the transposing loads fold into the LDS immediates,
// which takes ~a quarter of the scale warps' instructions away - they are what paces the kernel
// at p = 256, ncu: issue slots 63 % busy, tensor pipe 47 %), or 0 = use `pitch_rt`
template <int SUB, int PITCH, int RM, int NU = 4>
__device__ __forceinline__ void scale_col4(uint32_t r0, uint32_t pitch_rt, bool ok, uint32_t dsm,
                                           uint32_t Sp, uint32_t t_addr, int c, int kb, int ks,
                                           uint32_t vsm, float& gacc) {
    const uint32_t pitch = PITCH ? (uint32_t)PITCH : pitch_rt;
    // r0: shared address of element (row 0, column c) of the raw stage, pitch: bytes per row;
    // dsm, Sp: shared addresses of the stage's d vector and of the S tile
    constexpr bool s_lo = SUB == 1, t_lo = SUB == 2;   // 3xTF32 sub-pass: which operand is the residual
    float x[NU][4];
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            x[u][i] = ok ? lds_f32(r0 + (uint32_t)(4 * (kb + u * ks) + i) * pitch) : 0.f;
    const uint32_t tile_off = (uint32_t)(c >> 7) * TILE_BYTES;
    const uint32_t row = (uint32_t)c & 127u;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int k4 = kb + u * ks;
        const float4 dv = lds_f32x4(dsm + 16u * (uint32_t)k4);
        uint4 a;
        a.x = tf32_part<RM>(x[u][0], s_lo);
        a.y = tf32_part<RM>(x[u][1], s_lo);
        a.z = tf32_part<RM>(x[u][2], s_lo);
        a.w = tf32_part<RM>(x[u][3], s_lo);
        if (ok) sts_u32x4(Sp + tile_off + kmajor_chunk_off(row, (uint32_t)k4), a);
        tmem_st_x4(t_addr + (uint32_t)(4 * k4), tf32_part<RM>(dv.x * x[u][0], t_lo),
                   tf32_part<RM>(dv.y * x[u][1], t_lo), tf32_part<RM>(dv.z * x[u][2], t_lo),
                   tf32_part<RM>(dv.w * x[u][3], t_lo));
        if (vsm) {  // X^T v rides along in fp32 (not TF32: it is the score of an IRLS step)
            const float4 vv = lds_f32x4(vsm + 16u * (uint32_t)k4);
            gacc = fmaf(vv.x, x[u][0], gacc);
            gacc = fmaf(vv.y, x[u][1], gacc);
            gacc = fmaf(vv.z, x[u][2], gacc);
            gacc = fmaf(vv.w, x[u][3], gacc);
        }
    }
}

// The same for a COLUMN-MAJOR X (dense_helpers-tmpl.cpp:266-308 has C and F variants too): the
// TMA box [32 rows x P columns] of the (P x n) row-major array lands K-major in the stage - one
// 128-byte row of 32 k-values per X column, 16-byte chunks XOR-swizzled by the TMA unit
// (SWIZZLE_128B) so that the 32 lanes of a warp, one column each, read a chunk without bank
// conflicts.  No transpose is needed: a lane reads its 4 consecutive k as one LDS.128.
template <int SUB, int RM, int NU = 4>
__device__ __forceinline__ void scale_col4_f(uint32_t R, bool ok, uint32_t dsm, uint32_t Sp,
                                             uint32_t t_addr, int c, int kb, int ks, uint32_t vsm,
                                             float& gacc) {
    constexpr bool s_lo = SUB == 1, t_lo = SUB == 2;
    const uint32_t tile_off = (uint32_t)(c >> 7) * TILE_BYTES;
    const uint32_t row = (uint32_t)c & 127u;
    float4 x[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u)
        x[u] = ok ? lds_f32x4(R + kmajor_chunk_off((uint32_t)c, (uint32_t)(kb + u * ks)))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int k4 = kb + u * ks;
        const float4 dv = lds_f32x4(dsm + 16u * (uint32_t)k4);
        uint4 a;
        a.x = tf32_part<RM>(x[u].x, s_lo);
        a.y = tf32_part<RM>(x[u].y, s_lo);
        a.z = tf32_part<RM>(x[u].z, s_lo);
        a.w = tf32_part<RM>(x[u].w, s_lo);
        if (ok) sts_u32x4(Sp + tile_off + kmajor_chunk_off(row, (uint32_t)k4), a);
        tmem_st_x4(t_addr + (uint32_t)(4 * k4), tf32_part<RM>(dv.x * x[u].x, t_lo),
                   tf32_part<RM>(dv.y * x[u].y, t_lo), tf32_part<RM>(dv.z * x[u].z, t_lo),
                   tf32_part<RM>(dv.w * x[u].w, t_lo));
        if (vsm) {
            const float4 vv = lds_f32x4(vsm + 16u * (uint32_t)k4);
            gacc = fmaf(vv.x, x[u].x, gacc);
            gacc = fmaf(vv.y, x[u].y, gacc);
            gacc = fmaf(vv.z, x[u].z, gacc);
            gacc = fmaf(vv.w, x[u].w, gacc);
        }
    }
}

// one operand slot of a raw stage: sub-pass SUB of the 3xTF32 scheme (0 = the plain TF32 pass).
// Who does what.  8 scale warps: the two warps of a TMEM lane quarter (h = 0, 1) split the 8 row
// chunks of a column (P <= 128: chunks h, h+2, h+4, h+6) or the two 128-column tiles (P > 128,
// all 8 chunks).  16 scale warps (h = 0..3): chunks h, h+4 (P <= 128) or tile h >> 1, chunks
// (h & 1) + 0, 2, 4, 6.
template <int SUB, int RM, int NSW>
__device__ __forceinline__ void scale_stage(bool f_order, int mtiles, uint32_t R, uint32_t col_off,
                                            uint32_t pitch, bool ok, uint32_t dsm, uint32_t Sp,
                                            uint32_t t_addr, int my_col, int h, uint32_t vsm,
                                            float& gacc) {
    if (NSW == 16) {
        if (f_order) {
            if (mtiles == 1)
                scale_col4_f<SUB, RM, 2>(R, ok, dsm, Sp, t_addr, my_col, h, 4, vsm, gacc);
            else
                scale_col4_f<SUB, RM, 4>(R, ok, dsm, Sp, t_addr, my_col, h & 1, 2, vsm, gacc);
        } else if (mtiles == 1) {
            if (pitch == 512)
                scale_col4<SUB, 512, RM, 2>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, h, 4, vsm, gacc);
            else
                scale_col4<SUB, 0, RM, 2>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, h, 4, vsm, gacc);
        } else if (pitch == 1024) {
            scale_col4<SUB, 1024, RM, 4>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, h & 1, 2, vsm, gacc);
        } else if (pitch == 512) {
            scale_col4<SUB, 512, RM, 4>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, h & 1, 2, vsm, gacc);
        } else {
            scale_col4<SUB, 0, RM, 4>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, h & 1, 2, vsm, gacc);
        }
        return;
    }
    if (f_order) {
        if (mtiles == 1) {
            scale_col4_f<SUB, RM>(R, ok, dsm, Sp, t_addr, my_col, h, 2, vsm, gacc);
        } else {
            scale_col4_f<SUB, RM>(R, ok, dsm, Sp, t_addr, my_col, 0, 1, vsm, gacc);
            scale_col4_f<SUB, RM>(R, ok, dsm, Sp, t_addr, my_col, 4, 1, vsm, gacc);
        }
    } else if (mtiles == 1) {
        if (pitch == 512)
            scale_col4<SUB, 512, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, h, 2, vsm, gacc);
        else
            scale_col4<SUB, 0, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, h, 2, vsm, gacc);
    } else if (pitch == 1024) {
        scale_col4<SUB, 1024, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, 0, 1, vsm, gacc);
        scale_col4<SUB, 1024, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, 4, 1, vsm, gacc);
    } else if (pitch == 512) {
        scale_col4<SUB, 512, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, 0, 1, vsm, gacc);
        scale_col4<SUB, 512, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, 4, 1, vsm, gacc);
    } else {
        scale_col4<SUB, 0, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, 0, 1, vsm, gacc);
        scale_col4<SUB, 0, RM>(R + col_off, pitch, ok, dsm, Sp, t_addr, my_col, 4, 1, vsm, gacc);
    }
}

// fp32 -> tf32 of the MMA operands (to_tf32<RM>): -1 = TABMAT_B200_TC_ROUND or the default
int g_tc_round = -1;
// longest accumulation chain in MMA steps (tm_set_tc_flush_steps): -1 = env / default, 0 = unbounded
int g_tc_flush_steps = -1;
// SMs the persistent kernel does not occupy (tm_set_tc_sm_reserve)
int g_tc_sm_reserve = 0;
static int tc_round_mode() {
    int rm = g_tc_round;
    if (rm < 0) {
        static const int env = getenv("TABMAT_B200_TC_ROUND") ? atoi(getenv("TABMAT_B200_TC_ROUND")) : 2;
        rm = env;
    }
    return rm < 0 || rm > 2 ? 2 : rm;
}

// tensor maps of one launch: the X tile, the weight vector d and the one-hot code vectors
// (1-d maps, 32 elements per stage; out-of-range rows read as zero)
struct TmapSet {
    CUtensorMap x;
    CUtensorMap d;
    CUtensorMap codes[8];
    CUtensorMap v;
    CUtensorMap xb;   // panel mode: the box of panel B (its own width)
};

// bring-up timeline: cycle stamps of CTA 0's first TL_ITERS iterations (prm.dbg only)
constexpr int TL_ITERS = 1024;
constexpr int TL_OFF = 65536 + 3 * 128 * 128;  // float offset of the int64 timeline in dbg
__device__ __forceinline__ void tl_stamp(const Params& prm, int it, int e) {
    if (prm.dbg && blockIdx.x == 0 && it < TL_ITERS)
        reinterpret_cast<long long*>(prm.dbg + TL_OFF)[(size_t)it * 8 + e] = clock64();
}

// ---------------------------------------------------------------------------------------
// min-blocks 2 only caps the registers (<= 102/thread) so that the L2-atomic-bound scatter
// kernels of a SplitMatrix sandwich can share the SM with this kernel's one resident CTA
// SCW = number of scatter warps appended after the scale warps (0: plain SYRK + one-hot kernel)
// NSUB = operand slots per raw stage: 1 (TF32) or 3 (3xTF32, tm_set_dense_f32_mode(3)); a
// template parameter because the scale warps are the throughput-critical part of the pipeline
// and the plain path must not pay for the residual arithmetic
template <int MIN_BLOCKS, int SCW, int NSUB, int NSW>
__global__ void __launch_bounds__(tc_threads(NSW, SCW), MIN_BLOCKS)
k_dense_syrk_tc(const __grid_constant__ TmapSet tmaps, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int SR = prm.stagesR;
    const int SB = prm.sb;
    const int P = prm.P;
    const uint32_t half_bytes = (uint32_t)prm.mtiles * TILE_BYTES;   // S of one slot
    const uint32_t oh_bytes = (uint32_t)prm.oh_groups * GROUP_BYTES;
    const uint32_t slot_bytes = half_bytes + oh_bytes;

    // smem: [operand ring: SB x (S | one-hot)] [R ring: SR x r_bytes] [barriers]
    uint8_t* Oper = base;
    uint8_t* Rring = Oper + (size_t)SB * slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Rring + (size_t)SR * prm.r_bytes);
    uint64_t* full = bars;                          // [MAX_STAGES] TMA -> scale warps
    uint64_t* emptyR = bars + MAX_STAGES;           // [MAX_STAGES] scale warps -> producer
    uint64_t* scaled = bars + 2 * MAX_STAGES;           // [SB] scale warps -> MMA
    uint64_t* emptyB = bars + 2 * MAX_STAGES + MAX_SB;  // [SB] MMA -> scale warps
    uint64_t* done = bars + 2 * MAX_STAGES + 2 * MAX_SB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // operand ring starts as zeros: X columns >= P of a tile are never written, and the
    // one-hot tiles are all-zero except for the ones set (and cleared again) per stage
    {
        uint4* z = reinterpret_cast<uint4*>(Oper);
        for (uint32_t i = threadIdx.x; i < SB * slot_bytes / 16; i += tc_threads(NSW, SCW))
            z[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < SR; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&emptyR[s], NSW + SCW);
        }
        for (int b = 0; b < SB; ++b) {
            mbar_init(&scaled[b], NSW);  // one arrive per scale warp
            mbar_init(&emptyB[b], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmaps.x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmaps.d) : "memory");
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
    const int syrk_tiles = prm.dual_acc ? 2 : prm.mtiles * (prm.mtiles + 1) / 2;
    const uint32_t oh_col0 = (uint32_t)syrk_tiles * 128;  // first TMEM column of the one-hot block
    const uint32_t t_col0 = oh_col0 + (uint32_t)prm.oh_groups * 32;  // T operand ring (SB x mtiles x 32)

    // row tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    int my_count = 0;
    if ((long long)blockIdx.x < prm.num_row_tiles)
        my_count = (int)((prm.num_row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (warp == 0) {
        // ===== TMA producer (one lane): per stage one box of X plus 128 bytes of d and of each
        // one-hot code vector, all landing in the stage and completing the same mbarrier - no
        // thread of this CTA ever has a plain global load in flight =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const int n_codes = prm.oh_ncat + (prm.sc_staged ? prm.sc_ncat : 0);
            const uint32_t tx = (uint32_t)(BK * (prm.panel ? prm.wa + prm.wb : P) * 4) +
                                128u * (1u + (uint32_t)n_codes + (prm.has_v ? 1u : 0u));
            for (int it = 0; it < my_count; ++it, ++s) {
                if (s == SR) {
                    s = 0;
                    ph ^= 1;
                }
                mbar_wait_t<(SCW > 0)>(&emptyR[s], ph ^ 1);
                tl_stamp(prm, it, 0);
                const long long k0 = (prm.tile0 + (long long)blockIdx.x + (long long)it * gridDim.x) * BK;
                uint8_t* stage = Rring + (size_t)s * prm.r_bytes;
                const uint32_t aux = smem_u32(stage) + (uint32_t)prm.aux_off;
                mbar_expect_tx(&full[s], tx);
                if (prm.panel) {   // one box per 128-column panel, at its global column offset
                    if (prm.f_order) {
                        tma_load_2d(stage, &tmaps.x, &full[s], (int)k0, prm.a0);
                        if (prm.wb) tma_load_2d(stage + TILE_BYTES, &tmaps.xb, &full[s], (int)k0, prm.b0);
                    } else {
                        tma_load_2d(stage, &tmaps.x, &full[s], prm.a0, (int)k0);
                        if (prm.wb) tma_load_2d(stage + TILE_BYTES, &tmaps.xb, &full[s], prm.b0, (int)k0);
                    }
                } else if (prm.f_order)
                    tma_load_2d(stage, &tmaps.x, &full[s], (int)k0, 0);   // (k, column) coordinates
                else
                    tma_load_2d(stage, &tmaps.x, &full[s], 0, (int)k0);
                tma_load_1d(aux, &tmaps.d, &full[s], (int)k0);
                for (int c = 0; c < n_codes; ++c)
                    tma_load_1d(aux + 128u * (uint32_t)(c + 1), &tmaps.codes[c], &full[s], (int)k0);
                if (prm.has_v) tma_load_1d(aux + 128u * 9u, &tmaps.v, &full[s], (int)k0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // One thread issues; the per-MMA instruction count matters (a single thread's issue rate
        // is what paces small MMAs), so the descriptors are precomputed and only advanced by
        // constants.  With one 128-column tile the S tile and the one-hot rows are contiguous in
        // shared memory and their accumulators contiguous in TMEM, so each 8-row K step is ONE
        // MMA chain of N <= 256 pieces over [S ; one-hot]:  D[:, 0:128] is the SYRK tile and
        // D[:, 128:] the one-hot block.  Without one-hot blocks the K steps alternate between two
        // accumulator tiles (summed in the epilogue) to halve the dependent-accumulate stalls.
        const uint32_t idesc = make_idesc(128, 128);
        const uint64_t desc0 = make_desc(smem_u32(Oper));      // slot 0, k step 0
        const uint32_t slot16 = slot_bytes >> 4;               // descriptor units (16 B)
        const int n_total = 128 + prm.oh_groups * 32;          // mtiles == 1: columns of [S ; O]
        const int n1 = n_total > 256 ? 256 : n_total;
        const int n2 = n_total - n1;
        const uint32_t idesc1 = make_idesc(128, n1), idesc2 = make_idesc(128, n2 > 0 ? n2 : 16);
        int b = 0, sub = 0;
        uint32_t phb = 0;
        const int total_it = my_count * NSUB;   // operand slots: NSUB per raw stage
        for (int it = 0; it < total_it; ++it, ++b) {
            if (b == SB) {
                b = 0;
                phb ^= 1;
            }
            mbar_wait_t<(SCW > 0)>(&scaled[b], phb);
            tcgen05_fence_after();
            if (elect_one()) {
                tl_stamp(prm, it, 4);
                const uint64_t dS = desc0 + (uint64_t)((uint32_t)b * slot16);
                const uint32_t Ta = tmem_base + t_col0 + (uint32_t)(b * prm.mtiles) * 32;
                // 3xTF32 sub-pass 1 multiplies hi(d x) with the RESIDUAL of X: the one-hot
                // columns (exact 0 / 1, no residual) must not be added a second time
                const bool skip_oh = NSUB > 1 && sub == 1;
                if (prm.mtiles == 1) {
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                        const uint64_t db = dS + (uint64_t)(ks * 2);   // +32 B per k step
                        if (prm.dual_acc) {
                            // even k steps -> tile 0, odd -> tile 1 (first use of each overwrites)
                            const uint32_t acc2 = (it > 0 || ks > 1) ? 1u : 0u;
                            tcgen05_mma_tf32_ts(tmem_base + (uint32_t)(ks & 1) * 128, Ta + ks * 8,
                                                db, idesc, acc2);
                        } else if (skip_oh) {
                            tcgen05_mma_tf32_ts(tmem_base, Ta + ks * 8, db, idesc, acc);
                        } else {
                            tcgen05_mma_tf32_ts(tmem_base, Ta + ks * 8, db, idesc1, acc);
                            if (n2 > 0)
                                tcgen05_mma_tf32_ts(tmem_base + 256, Ta + ks * 8,
                                                    db + (uint64_t)(256 * 128 / 16), idesc2, acc);
                        }
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                        const uint64_t d0 = dS + (uint64_t)(ks * 2);
                        const uint64_t d1 = d0 + (uint64_t)(TILE_BYTES / 16);
                        if (prm.tile_mask & 1)
                            tcgen05_mma_tf32_ts(tmem_base, Ta + ks * 8, d0, idesc, acc);             // (0,0)
                        if (prm.tile_mask & 2)
                            tcgen05_mma_tf32_ts(tmem_base + 128, Ta + 32 + ks * 8, d0, idesc, acc);  // (1,0)
                        if (prm.tile_mask & 4)
                            tcgen05_mma_tf32_ts(tmem_base + 256, Ta + 32 + ks * 8, d1, idesc, acc);  // (1,1)
                    }
                }
                tcgen05_commit(&emptyB[b]);
                tl_stamp(prm, it, 5);
            }
            __syncwarp();
            if (++sub == NSUB) sub = 0;
        }
        if (elect_one()) tcgen05_commit(done);
        __syncwarp();
    } else if (SCW > 0 && warp >= 2 + NSW) {
        // ===== scatter warps: vector REDs of the cross blocks that share the dense operand =====
        // Warp ws owns rows [ws * RPW, (ws + 1) * RPW) of every row tile of this CTA.  It copies
        // its rows from the raw stage into registers (one LDS.128 per row: lane l holds columns
        // 4l..4l+3), releases the stage, and then issues
        //   out_sparse[j, :]      += A[k, j] * d[k] * X[k, :]     per non-zero (k, j)
        //   tab_c[code_c[k], :]   += d[k] * X[k, :]               summed in registers over runs of
        //                                                          equal codes (row-sorted storage)
        // The CSR slice and the codes of a tile are prefetched one tile ahead (indptr two tiles
        // ahead) with plain loads into registers, lane-parallel, and broadcast with shuffles.
        constexpr int RPW = SCW > 0 ? BK / SCW : 1;       // rows per warp and tile (8 or 4)
        constexpr int NCM = TC_SCATTER_MAX_CATS;
        const unsigned FULL = 0xffffffffu;
        const int ws = warp - (2 + NSW);
        const int r0 = ws * RPW;
        const bool lane_ok = lane * 4 < P;
        const int nc = prm.sc_ncat;
        const bool has_sp = prm.out_sparse != nullptr;
        const uint32_t rring_sa = smem_u32(Rring);
        if (prm.sc_staged && !has_sp) {
            // ---- categorical blocks only, codes staged by TMA: no global load anywhere in this
            // warp.  A lone warp runs at one dependent instruction per ~5 cycles (ncu: the first
            // form of this path executed ~680 branchy instructions per tile and needed ~6400
            // cycles for them, 4x the MMA pipeline's tile period), so the code per tile is kept
            // short and straight: all lanes read the same d / code words (broadcast LDS.128), the
            // outer loop over the blocks skips the unused ones, the run logic is one warp-uniform
            // compare per row, the adds are predicated FMAs.  A flush copies the sum to a second
            // register set and REDs from there: the RED holds its source registers for hundreds
            // of cycles, the accumulator itself is free at once.
            float* tabs[NCM];
            float4 accs[NCM], outb[NCM];
            int curc[NCM];
#pragma unroll
            for (int c = 0; c < NCM; ++c) {
                curc[c] = -1;
                accs[c] = outb[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                tabs[c] = nullptr;
                if (c < nc)
                    tabs[c] = prm.sc_tab[c] +
                              (size_t)((blockIdx.x * SCW + ws) % prm.sc_copies[c]) * (size_t)prm.sc_K[c] * P +
                              lane * 4;
            }
            int s2 = 0;
            uint32_t ph2 = 0;
            for (int it = 0; it < my_count; ++it, ++s2) {
                if (s2 == SR) {
                    s2 = 0;
                    ph2 ^= 1;
                }
                if (lane == 0) mbar_wait_sleep(&full[s2], ph2);
                __syncwarp();
                const uint32_t stage = rring_sa + (uint32_t)s2 * (uint32_t)prm.r_bytes;
                const uint32_t aux = stage + (uint32_t)prm.aux_off;
                float4 y[RPW];
                float dq[RPW];
#pragma unroll
                for (int q = 0; q < RPW; ++q)
                    y[q] = lane_ok ? lds_f32x4(stage + (uint32_t)(r0 + q) * (uint32_t)P * 4u + (uint32_t)lane * 16u)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q4 = 0; q4 < RPW; q4 += 4) {
                    const float4 dd = lds_f32x4(aux + (uint32_t)(r0 + q4) * 4u);
                    dq[q4] = dd.x, dq[q4 + 1] = dd.y, dq[q4 + 2] = dd.z, dq[q4 + 3] = dd.w;
                }
                int codes[NCM][RPW];
                uint32_t dep = 0;   // depends on every value loaded from the stage
#pragma unroll
                for (int c = 0; c < NCM; ++c) {
                    const uint32_t ca = aux + 128u * (uint32_t)(1 + prm.oh_ncat + c) + (uint32_t)r0 * 4u;
#pragma unroll
                    for (int q4 = 0; q4 < RPW; q4 += 4) {
                        if (c < nc) {
                            const float4 cc = lds_f32x4(ca + (uint32_t)q4 * 4u);   // 4 int32 codes
                            codes[c][q4] = __float_as_int(cc.x), codes[c][q4 + 1] = __float_as_int(cc.y);
                            codes[c][q4 + 2] = __float_as_int(cc.z), codes[c][q4 + 3] = __float_as_int(cc.w);
                            dep ^= (uint32_t)codes[c][q4 + 3];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < RPW; ++q) dep ^= __float_as_uint(y[q].w) ^ __float_as_uint(dq[q]);
                // all reads of the stage are in registers: hand it back at once (relaxed: the REDs
                // of earlier tiles may still be in flight; the dependency keeps the arrive behind
                // the loads)
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed_after(&emptyR[s2], dep);
#pragma unroll
                for (int c = 0; c < NCM; ++c) {
                    if (c >= nc) continue;   // warp-uniform: unused blocks cost nothing
                    int (&code)[RPW] = codes[c];
                    const int df = prm.sc_df[c];
                    bool same = true;   // the whole slice carries one code (row-sorted storage)
#pragma unroll
                    for (int q = 1; q < RPW; ++q) same = same && code[q] == code[0];
                    if (same) {
                        // fast path: at most one flush, then 4 * RPW straight FMAs (rows with
                        // weight 0 add 0)
                        int cd = code[0] - df;
                        cd = cd < 0 ? -1 : cd;
                        if (cd != curc[c]) {
                            outb[c] = accs[c];
                            if (curc[c] >= 0 && lane_ok) red_add_v4(tabs[c] + (size_t)curc[c] * P, outb[c]);
                            accs[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                            curc[c] = cd;
                        }
                        if (cd >= 0) {
#pragma unroll
                            for (int q = 0; q < RPW; ++q) {
                                accs[c].x = fmaf(y[q].x, dq[q], accs[c].x);
                                accs[c].y = fmaf(y[q].y, dq[q], accs[c].y);
                                accs[c].z = fmaf(y[q].z, dq[q], accs[c].z);
                                accs[c].w = fmaf(y[q].w, dq[q], accs[c].w);
                            }
                        }
                        continue;
                    }
#pragma unroll
                    for (int q = 0; q < RPW; ++q) {
                        const float dk = dq[q];
                        int cd = code[q] - df;
                        cd = cd < 0 ? -1 : cd;
                        cd = dk == 0.f ? curc[c] : cd;   // a zero-weight row never ends a run
                        if (cd != curc[c]) {             // warp-uniform, rare on sorted rows
                            outb[c] = accs[c];
                            if (curc[c] >= 0 && lane_ok) red_add_v4(tabs[c] + (size_t)curc[c] * P, outb[c]);
                            accs[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                            curc[c] = cd;
                        }
                        if (cd >= 0) {
                            accs[c].x = fmaf(y[q].x, dk, accs[c].x);
                            accs[c].y = fmaf(y[q].y, dk, accs[c].y);
                            accs[c].z = fmaf(y[q].z, dk, accs[c].z);
                            accs[c].w = fmaf(y[q].w, dk, accs[c].w);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NCM; ++c)
                if (c < nc && curc[c] >= 0 && lane_ok) red_add_v4(tabs[c] + (size_t)curc[c] * P, accs[c]);
        } else {
        float* tab[NCM];
        float4 acc[NCM];
        int cur[NCM];
#pragma unroll
        for (int c = 0; c < NCM; ++c) {
            cur[c] = -1;
            acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            tab[c] = nullptr;
            if (c < nc)
                tab[c] = prm.sc_tab[c] +
                         (size_t)((blockIdx.x * SCW + ws) % prm.sc_copies[c]) * (size_t)prm.sc_K[c] * P +
                         lane * 4;
        }
        float* const osp = prm.out_sparse + lane * 4;
        auto tile_row0 = [&](int it) -> long long {
            return (prm.tile0 + (long long)blockIdx.x + (long long)it * gridDim.x) * BK + r0;
        };
        // indptr of rows k .. k + RPW (lanes 0..RPW), clamped to n (rows past the end are empty)
        auto load_ip = [&](int it) -> int {
            if (!has_sp || it >= my_count || lane > RPW) return 0;
            long long k = tile_row0(it) + lane;
            if (k > prm.n) k = prm.n;
            return __ldg(prm.csr_indptr + k);
        };
        int ip_cur = load_ip(0), ip_nxt = load_ip(1);
        int idx_cur = 0, idx_nxt = 0;
        float val_cur = 0.f, val_nxt = 0.f;
        int code_cur[NCM], code_nxt[NCM];
#pragma unroll
        for (int c = 0; c < NCM; ++c) code_cur[c] = code_nxt[c] = -1;
        auto load_entries = [&](int it, int ip, int& idx, float& val, int (&code)[NCM]) {
            idx = 0;
            val = 0.f;
#pragma unroll
            for (int c = 0; c < NCM; ++c) code[c] = -1;
            if (it >= my_count) return;
            if (has_sp) {
                const int e0 = __shfl_sync(FULL, ip, 0), e1 = __shfl_sync(FULL, ip, RPW);
                if (e0 + lane < e1) {
                    idx = __ldg(prm.csr_indices + e0 + lane);
                    val = __ldg(prm.csr_data + e0 + lane);
                }
            }
            const long long k = tile_row0(it) + lane;
            if (lane < RPW && k < prm.n) {
#pragma unroll
                for (int c = 0; c < NCM; ++c)
                    if (c < nc) {
                        const int v = __ldg(prm.sc_codes[c] + k) - prm.sc_df[c];
                        code[c] = v < 0 ? -1 : v;
                    }
            }
        };
        load_entries(0, ip_cur, idx_cur, val_cur, code_cur);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < my_count; ++it, ++s) {
            if (s == SR) {
                s = 0;
                ph ^= 1;
            }
            // prefetch: entries of tile it + 1 (its indptr was requested a tile ago), indptr of it + 2
            load_entries(it + 1, ip_nxt, idx_nxt, val_nxt, code_nxt);
            const int ip_nn = load_ip(it + 2);
            if (lane == 0) mbar_wait_sleep(&full[s], ph);
            __syncwarp();
            const uint32_t stage = rring_sa + (uint32_t)s * (uint32_t)prm.r_bytes;
            float4 y[RPW];
            float dq[RPW];
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
                y[q] = lane_ok ? lds_f32x4(stage + (uint32_t)(r0 + q) * (uint32_t)P * 4u + (uint32_t)lane * 16u)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
                dq[q] = lds_f32(stage + (uint32_t)prm.aux_off + (uint32_t)(r0 + q) * 4u);
            }
            const int ebase = __shfl_sync(FULL, ip_cur, 0);
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
                const float dk = dq[q];
                int e0 = __shfl_sync(FULL, ip_cur, q), e1 = __shfl_sync(FULL, ip_cur, q + 1);
                int cq[NCM];
#pragma unroll
                for (int c = 0; c < NCM; ++c) cq[c] = __shfl_sync(FULL, code_cur[c], q);
                if (dk == 0.f) continue;  // every term of row k is proportional to d[k]
                const float4 yy = make_float4(y[q].x * dk, y[q].y * dk, y[q].z * dk, y[q].w * dk);
#pragma unroll
                for (int c = 0; c < NCM; ++c) {
                    if (c >= nc) continue;
                    if (cq[c] != cur[c]) {  // warp-uniform: a run of block c ends here
                        if (cur[c] >= 0 && lane_ok) red_add_v4(tab[c] + (size_t)cur[c] * P, acc[c]);
                        acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                        cur[c] = cq[c];
                    }
                    if (cq[c] >= 0) {
                        acc[c].x += yy.x;
                        acc[c].y += yy.y;
                        acc[c].z += yy.z;
                        acc[c].w += yy.w;
                    }
                }
                // (groups of 4 REDs with independent value registers were measured: no gain — a
                // warp's REDs are serialised by the memory pipeline, ~500 cycles each on B200)
                for (int e = e0; e < e1; ++e) {
                    const int off = e - ebase;
                    int j;
                    float a;
                    if (off < 32) {  // warp-uniform
                        j = __shfl_sync(FULL, idx_cur, off);
                        a = __shfl_sync(FULL, val_cur, off);
                    } else {         // more than 32 non-zeros in RPW rows: straight from memory
                        j = __ldg(prm.csr_indices + e);
                        a = __ldg(prm.csr_data + e);
                    }
                    if (lane_ok)
                        red_add_v4(osp + (size_t)j * P,
                                   make_float4(yy.x * a, yy.y * a, yy.z * a, yy.w * a));
                }
            }
            // the rows read from the stage have been consumed: hand it back (relaxed arrive: do
            // not wait for the REDs above)
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(&emptyR[s]);
            ip_cur = ip_nxt;
            ip_nxt = ip_nn;
            idx_cur = idx_nxt;
            val_cur = val_nxt;
#pragma unroll
            for (int c = 0; c < NCM; ++c) code_cur[c] = code_nxt[c];
        }
#pragma unroll
        for (int c = 0; c < NCM; ++c)
            if (c < nc && cur[c] >= 0 && lane_ok) red_add_v4(tab[c] + (size_t)cur[c] * P, acc[c]);
        }   // general (sparse + global-load) path
    } else {
        // ===== scale warps, then epilogue =====
        const int w = warp - 2;                // 0..NSW-1
        const int t = (int)threadIdx.x - 64;   // 0..32*NSW-1
        // one-hot: thread t < 256 encodes row (t >> 3) of categorical block (t & 7)
        const int oh_r = (t >> 3) & 31;
        const int oh_c = t & 7;
        const bool oh_thread = oh_c < prm.oh_ncat && t < 256;
        const int my_off = oh_thread ? prm.oh_off[oh_c] : 0;
        const int my_K = oh_thread ? prm.oh_K[oh_c] : 0;
        const int my_df = oh_thread ? prm.oh_df[oh_c] : 0;
        uint32_t prev0 = 0xffffffffu, prev1 = 0xffffffffu, prev2 = 0xffffffffu,
                 prev3 = 0xffffffffu;  // the one this thread set in operand slot 0..3
        // X column of this thread: TMEM lane quarter q = warp & 3 (a warp can only touch its own
        // quarter); the two warps of a quarter split the 8 row chunks (P <= 128) or the two
        // 128-column tiles (P > 128)
        const int q = warp & 3;
        const int h = w >> 2;
        const int my_col = (prm.mtiles == 2 ? (NSW == 16 ? h >> 1 : h) * 128 : 0) + q * 32 + lane;
        const uint32_t oper_sa = smem_u32(Oper), rring_sa = smem_u32(Rring);
        // where this thread's column sits in a raw stage (row-major X)
        uint32_t col_off, col_pitch;
        bool col_ok;
        if (prm.panel) {
            const bool in_b = my_col >= 128;
            col_ok = in_b ? (my_col - 128) < prm.wb : my_col < prm.wa;
            col_pitch = (uint32_t)(in_b ? prm.wb : prm.wa) * 4u;
            col_off = in_b ? (uint32_t)TILE_BYTES + (uint32_t)(my_col - 128) * 4u : (uint32_t)my_col * 4u;
        } else {
            col_ok = my_col < P;
            col_pitch = (uint32_t)P * 4u;
            col_off = (uint32_t)my_col * 4u;
        }
        float gacc = 0.f;   // this thread's share of (X^T v)[my_col]
        int s = 0, b = 0;
        uint32_t ph = 0, phb = 0;
        for (int it = 0; it < my_count; ++it, ++s) {
            if (s == SR) {
                s = 0;
                ph ^= 1;
            }
            if (lane == 0) mbar_wait_t<(SCW > 0)>(&full[s], ph);
            __syncwarp();
            if (t == 0) tl_stamp(prm, it, 2);
          for (int sub = 0; sub < NSUB; ++sub, ++b) {   // 3xTF32: three operand slots per stage
            if (b == SB) {
                b = 0;
                phb ^= 1;
            }
            // one lane polls, the warp follows through __syncwarp
            if (lane == 0) mbar_wait_t<(SCW > 0)>(&emptyB[b], phb ^ 1);  // MMAs of slot use it-SB left slot b
            __syncwarp();
            if (t == 0) tl_stamp(prm, it, 1);
            const uint32_t Sp = oper_sa + (uint32_t)b * slot_bytes;
            const uint32_t stage = rring_sa + (uint32_t)s * (uint32_t)prm.r_bytes;
            const uint32_t dsm = stage + (uint32_t)prm.aux_off;
            if (prm.oh_groups) {
                const uint32_t O = Sp + half_bytes;
                const uint32_t pa = b == 0 ? prev0 : (b == 1 ? prev1 : (b == 2 ? prev2 : prev3));
                if (pa != 0xffffffffu) sts_f32(O + pa, 0.f);
                uint32_t na = 0xffffffffu;
                if (oh_thread) {
                    const int code = lds_s32(dsm + 128u + (uint32_t)(oh_c * 32 + oh_r) * 4u) - my_df;
                    if (code >= 0 && code < my_K && lds_f32(dsm + (uint32_t)oh_r * 4u) != 0.f) {
                        const uint32_t slot = (uint32_t)(my_off + code);
                        na = kmajor_chunk_off(slot, (uint32_t)oh_r >> 2) +
                             (((uint32_t)oh_r & 3u) << 2);
                        sts_f32(O + na, 1.0f);
                    }
                }
                if (b == 0) prev0 = na;
                else if (b == 1) prev1 = na;
                else if (b == 2) prev2 = na;
                else prev3 = na;
            }
            const uint32_t R = stage;
            {
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + t_col0 +
                                        (uint32_t)(b * prm.mtiles + (my_col >> 7)) * 32;
                const uint32_t vsm = (prm.has_v && sub == 0) ? dsm + 128u * 9u : 0u;
                if (NSUB == 1) {
                    if (prm.round_mode == 2)
                        scale_stage<0, 2, NSW>(prm.f_order, prm.mtiles, R, col_off, col_pitch, col_ok,
                                          dsm, Sp, t_addr, my_col, h, vsm, gacc);
                    else if (prm.round_mode == 1)
                        scale_stage<0, 1, NSW>(prm.f_order, prm.mtiles, R, col_off, col_pitch, col_ok,
                                          dsm, Sp, t_addr, my_col, h, vsm, gacc);
                    else
                        scale_stage<0, 0, NSW>(prm.f_order, prm.mtiles, R, col_off, col_pitch, col_ok,
                                          dsm, Sp, t_addr, my_col, h, vsm, gacc);
                } else if (sub == 0)
                    scale_stage<0, 1, NSW>(prm.f_order, prm.mtiles, R, col_off, col_pitch, col_ok, dsm,
                                      Sp, t_addr, my_col, h, vsm, gacc);
                else if (sub == 1)
                    scale_stage<1, 1, NSW>(prm.f_order, prm.mtiles, R, col_off, col_pitch, col_ok, dsm,
                                      Sp, t_addr, my_col, h, 0u, gacc);
                else
                    scale_stage<2, 1, NSW>(prm.f_order, prm.mtiles, R, col_off, col_pitch, col_ok, dsm,
                                      Sp, t_addr, my_col, h, 0u, gacc);
            }
            if (t == 0) tl_stamp(prm, it, 6);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tcgen05_fence_before();
            if (t == 0) tl_stamp(prm, it, 7);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (sub == NSUB - 1) mbar_arrive(&emptyR[s]);   // the raw stage can be refilled
                mbar_arrive(&scaled[b]);   // the operands are ready for the MMA warp
            }
            if (t == 0) tl_stamp(prm, it, 3);
          }
        }

        if (prm.has_v && my_col < P && my_count > 0) atomicAdd(&prm.vec_out[my_col], gacc);
        // epilogue: TMEM -> registers -> RED into `out` (transposed: lanes = output columns)
        mbar_wait_t<(SCW > 0)>(done, 0);
        tcgen05_fence_after();
        const int chalf = (warp - 2) >> 2;    // which share of the column chunks this warp drains
        constexpr int CPW = 16 / NSW;         // 32-column chunks of a 128-column tile per warp
        if (my_count > 0) {
            // local column -> global column of the result (legacy: identity)
            auto gcol = [&](int c) -> int {
                if (!prm.panel) return c < P ? c : -1;
                if (c < 128) return c < prm.wa ? prm.a0 + c : -1;
                return (c - 128) < prm.wb ? prm.b0 + (c - 128) : -1;
            };
            const int ldo = prm.panel ? prm.ldo : P;
            int tile = 0;
            for (int mt = 0; mt < prm.mtiles; ++mt) {
                for (int nt = 0; nt <= mt; ++nt, ++tile) {
                    if (prm.mtiles == 2 && !((prm.tile_mask >> tile) & 1)) continue;
                    const int C = gcol(mt * 128 + q * 32 + lane);  // output column (= X column of A)
#pragma unroll 1
                    for (int cc = 0; cc < (prm.dual_acc ? 2 : 1) * CPW; ++cc) {
                        // dual_acc: the second accumulator tile (TMEM columns 128..255) holds the
                        // odd k steps of the same output tile
                        const int dup = cc / CPW;
                        const int n0 = chalf * (32 * CPW) + (cc % CPW) * 32;
                        uint32_t v[32];
                        uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                         (uint32_t)((tile + dup) * 128 + n0);
                        TM_TMEM_LD_32x32B_X32(taddr, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (prm.dbg && blockIdx.x == 0) {
                            float* T = prm.dbg + 65536 + (size_t)tile * 128 * 128;
                            for (int j = 0; j < 32; ++j)
                                T[(q * 32 + lane) * 128 + n0 + j] = __uint_as_float(v[j]);
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int Rr = gcol(nt * 128 + n0 + j);  // output row (= X column of B)
                            if (Rr >= 0 && C >= Rr)
                                atomicAdd(&prm.out[(size_t)Rr * ldo + C], __uint_as_float(v[j]));
                        }
                    }
                }
            }
            // one-hot block: D'[col = lane, slot] -> oh_out[slot, col]
            const int C = q * 32 + lane;
#pragma unroll 1
            for (int ch = chalf; ch < prm.oh_groups; ch += NSW / 4) {
                uint32_t v[32];
                uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + oh_col0 + (uint32_t)ch * 32;
                TM_TMEM_LD_32x32B_X32(taddr, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int slot = ch * 32 + j;
                    if (slot < prm.oh_slots && C < P) {
                        float val = __uint_as_float(v[j]);
                        if (val != 0.f) atomicAdd(&prm.oh_out[(size_t)slot * P + C], val);
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
    // p > 256 runs as passes over pairs of 128-column panels (TABMAT_B200_TC_MAX_P caps it)
    static const int64_t max_p = [] {
        const char* e = getenv("TABMAT_B200_TC_MAX_P");
        int64_t v = e ? atoll(e) : 2048;
        return v < 8 ? 8 : v;
    }();
    if (p < 8 || p > max_p) return false;
    // TMA: the pitch of the outer dimension must be a multiple of 16 bytes
    if (c_order ? (p % 4) != 0 : (n % 4) != 0) return false;
    static const bool f_off =
        getenv("TABMAT_B200_TC_FORDER") && atoi(getenv("TABMAT_B200_TC_FORDER")) == 0;
    if (!c_order && f_off) return false;
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0) return false;
    if (n < 1 || n > 0x7fffffffLL) return false;
    if (tc::device_cc_major() != 10) return false;
    return tc::get_encode() != nullptr;
}

float* g_tc_dbg = nullptr;
int g_tc_variant = 0;

// scatter warps of the fused form: TABMAT_B200_TC_SCW = 0 (default: separate scatter pass), 4 or 8.
// Measured on B200 at the benchmark shape (profiles/bench_r2a_*): with 4 / 8 scatter warps per SM
// the fused kernel is LATENCY-bound on its REDs (one RED per ~500 cycles and warp: 52 / 28 ms
// against 6.6 + 11.5 ms for the two separate passes, whose 64 warps per SM keep the L2 atomic
// units busy), so the separate scatter pass stays the default
int g_tc_scatter_warps = [] {
    const char* e = getenv("TABMAT_B200_TC_SCW");
    int v = e ? atoi(e) : -1;
    return (v == 0 || v == 4 || v == 8) ? v : -1;
}();
// -1 = auto (= none, see below); 4 / 8 force the fused form
static int tc_scatter_warps(bool with_sparse) {
    // auto: none when the per-non-zero REDs of dense x sparse would ride along (latency-bound,
    // above).  With only the run-aggregated categorical REDs left (dense x sparse computed by
    // the gather kernel) the first form still ran 30 ms instead of 6.6 (profiles/bench_r2c_*):
    // its per-tile global loads of the codes sat on the critical path of the raw-stage ring; the
    // codes are now staged by TMA (Params::sc_staged) and TABMAT_B200_TC_SCW_CATS picks the warp
    // count for that form (default 4)
    if (g_tc_scatter_warps >= 0) return g_tc_scatter_warps;
    if (with_sparse) return 0;
    static const int cats = [] {
        const char* e = getenv("TABMAT_B200_TC_SCW_CATS");
        int v = e ? atoi(e) : 4;
        return (v == 0 || v == 4 || v == 8) ? v : 4;
    }();
    return cats;
}
bool dense_tc_scatter_eligible(int64_t p, int n_cat, bool with_sparse) {
    return tc_scatter_warps(with_sparse) > 0 && p <= 128 && n_cat <= TC_SCATTER_MAX_CATS;
}

template <int MB, int SCW, int NSUB, int NSW = 8>
static int launch_tc(const tc::TmapSet& tmaps, const tc::Params& prm, unsigned grid, size_t smem,
                     cudaStream_t st) {
    // per device and cheap: set on every launch rather than cached in a process-wide static
    TM_CUDA(cudaFuncSetAttribute(tc::k_dense_syrk_tc<MB, SCW, NSUB, NSW>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    tc::k_dense_syrk_tc<MB, SCW, NSUB, NSW><<<grid, tc::tc_threads(NSW, SCW), smem, st>>>(tmaps, prm);
    TM_LAUNCHED();
    return 0;
}

// one launch on two 128-column panels of a wider matrix (p > 256); see tc::Params::panel
struct PanelSpec {
    int a0, wa, b0, wb, mask;
    int64_t p_full;
};

static int dense_sandwich_tc_launch(const float* X, int64_t n, int64_t p, int c_order, const float* d,
                                    float* out, cudaStream_t st, const TcOneHot* oh, bool share_sm,
                                    const FusedCrossParams* scatter, const float* v, float* vec_out,
                                    const PanelSpec* ps);

int dense_sandwich_tc_f32(const float* X, int64_t n, int64_t p, int c_order, const float* d,
                          float* out, cudaStream_t st, const TcOneHot* oh, bool share_sm,
                          const FusedCrossParams* scatter, const float* v, float* vec_out) {
    if (p <= 256)
        return dense_sandwich_tc_launch(X, n, p, c_order, d, out, st, oh, share_sm, scatter, v,
                                        vec_out, nullptr);
    // p > 256: 128-column panels.  Panel pairs (2m, 2m+1) give the three tiles of their 256 x 256
    // diagonal block in one pass; every remaining cross tile (I, J), J < I, is one pass over the
    // two panels that issues only that tile's MMAs.  Every pass streams 256 columns of X, so the
    // traffic grows with the number of passes (6 for p = 512, 28 for p = 1024) - still an order
    // of magnitude below the CUDA-core kernel, whose cost grows with p^2.
    if (oh || scatter || v) return fail("dense_tc: p > 256 is a plain SYRK");
    TM_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(p * p), st));
    const int np = (int)((p + 127) / 128);
    auto width = [&](int I) { return (int)(I == np - 1 ? p - (int64_t)I * 128 : 128); };
    for (int m = 0; 2 * m < np; ++m) {
        const int A = 2 * m, B = 2 * m + 1;
        PanelSpec ps{A * 128, width(A), B < np ? B * 128 : 0, B < np ? width(B) : 0, 7, p};
        int rc = dense_sandwich_tc_launch(X, n, p, c_order, d, out, st, nullptr, false, nullptr,
                                          nullptr, nullptr, &ps);
        if (rc) return rc;
    }
    for (int I = 0; I < np; ++I)
        for (int J = 0; J < I; ++J) {
            if (J == I - 1 && (I & 1)) continue;   // inside a diagonal pair
            PanelSpec ps{J * 128, width(J), I * 128, width(I), 2, p};
            int rc = dense_sandwich_tc_launch(X, n, p, c_order, d, out, st, nullptr, false, nullptr,
                                              nullptr, nullptr, &ps);
            if (rc) return rc;
        }
    return symmetrize_from_upper<float>(out, p, st);
}

static int dense_sandwich_tc_launch(const float* X, int64_t n, int64_t p, int c_order, const float* d,
                                    float* out, cudaStream_t st, const TcOneHot* oh, bool share_sm,
                                    const FusedCrossParams* scatter, const float* v, float* vec_out,
                                    const PanelSpec* ps) {
    using namespace tc;
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail("cuTensorMapEncodeTiled not available");
    if (!c_order && (oh || scatter))
        return fail("dense_tc: one-hot / scatter work needs a row-major dense block");

    TmapSet tmaps;
    memset(&tmaps, 0, sizeof(tmaps));
    const int64_t p_full = ps ? ps->p_full : p;
    if (ps) {
        p = ps->wb ? 128 + ps->wb : ps->wa;   // the kernel's local column count
        auto enc_box = [&](CUtensorMap* m, int w) -> bool {
            if (!c_order) {
                cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)p_full};
                cuuint64_t gstride[1] = {(cuuint64_t)n * sizeof(float)};
                cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)w};
                cuuint32_t estr[2] = {1, 1};
                return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstride,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            cuuint64_t gdim[2] = {(cuuint64_t)p_full, (cuuint64_t)n};
            cuuint64_t gstride[1] = {(cuuint64_t)p_full * sizeof(float)};
            cuuint32_t box[2] = {(cuuint32_t)w, (cuuint32_t)BK};
            cuuint32_t estr[2] = {1, 1};
            return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstride, box,
                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        };
        if (!enc_box(&tmaps.x, ps->wa)) return fail("cuTensorMapEncodeTiled failed (panel A)");
        if (ps->wb && !enc_box(&tmaps.xb, ps->wb)) return fail("cuTensorMapEncodeTiled failed (panel B)");
    } else if (!c_order) {
        // column-major X = (p x n) row-major: box of 32 consecutive rows (inner) x p columns
        cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)p};
        cuuint64_t gstride[1] = {(cuuint64_t)n * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)p};
        cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&tmaps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim,
                          gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (X, column-major)");
    } else {
        cuuint64_t gdim[2] = {(cuuint64_t)p, (cuuint64_t)n};
        cuuint64_t gstride[1] = {(cuuint64_t)p * sizeof(float)};
        cuuint32_t box[2] = {(cuuint32_t)p, (cuuint32_t)BK};  // one box = the whole row tile
        cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&tmaps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim,
                          gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (X)");
    }
    auto encode_1d = [&](CUtensorMap* m, const void* ptr, CUtensorMapDataType dt) -> bool {
        cuuint64_t gdim[1] = {(cuuint64_t)n};
        cuuint64_t gstride[1] = {0};  // unused for rank 1
        cuuint32_t box[1] = {(cuuint32_t)BK};
        cuuint32_t estr[1] = {1};
        return enc(m, dt, 1, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    // TMA needs 16-byte aligned global addresses: a misaligned d is staged through scratch
    Scratch d_al(((reinterpret_cast<uintptr_t>(d) & 15) != 0) ? sizeof(float) * (size_t)n : 0, st);
    if (d_al.err != cudaSuccess) return fail_cuda(d_al.err, "scratch");
    if ((reinterpret_cast<uintptr_t>(d) & 15) != 0) {
        TM_CUDA(cudaMemcpyAsync(d_al.p, d, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        d = d_al.as<float>();
    }
    if (!encode_1d(&tmaps.d, d, CU_TENSOR_MAP_DATA_TYPE_FLOAT32))
        return fail("cuTensorMapEncodeTiled failed (d)");
    Scratch v_al((v && (reinterpret_cast<uintptr_t>(v) & 15) != 0) ? sizeof(float) * (size_t)n : 0, st);
    if (v_al.err != cudaSuccess) return fail_cuda(v_al.err, "scratch");
    if (v) {
        if (!vec_out) return fail("dense_tc: v without vec_out");
        if ((reinterpret_cast<uintptr_t>(v) & 15) != 0) {
            TM_CUDA(cudaMemcpyAsync(v_al.p, v, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
            v = v_al.as<float>();
        }
        if (!encode_1d(&tmaps.v, v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32))
            return fail("cuTensorMapEncodeTiled failed (v)");
        TM_CUDA(cudaMemsetAsync(vec_out, 0, sizeof(float) * (size_t)p, st));
    }
    if (oh) {
        for (int c = 0; c < oh->ncat && c < 8; ++c) {
            if ((reinterpret_cast<uintptr_t>(oh->codes[c]) & 15) != 0)
                return fail("dense_tc: categorical code vectors must be 16-byte aligned");
            if (!encode_1d(&tmaps.codes[c], oh->codes[c], CU_TENSOR_MAP_DATA_TYPE_INT32))
                return fail("cuTensorMapEncodeTiled failed (codes)");
        }
    }

    Params prm;
    memset(&prm, 0, sizeof(prm));
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
    prm.dual_acc = (prm.mtiles == 1 && !(oh && oh->ncat > 0)) ? 1 : 0;
    if (prm.dual_acc) ntiles = 2;
    if (oh && oh->ncat > 0) {
        if (prm.mtiles != 1) return fail("dense_tc: one-hot blocks need p <= 128");
        if (oh->ncat > 8) return fail("dense_tc: at most 8 one-hot blocks");
        int slots = 0;
        for (int c = 0; c < oh->ncat; ++c) {
            prm.oh_codes[c] = oh->codes[c];
            prm.oh_off[c] = slots;
            prm.oh_K[c] = oh->K[c];
            prm.oh_df[c] = oh->drop_first[c];
            slots += oh->K[c];
        }
        if (slots > TC_ONEHOT_MAX_SLOTS) return fail("dense_tc: too many one-hot slots");
        prm.oh_ncat = oh->ncat;
        prm.oh_slots = slots;
        prm.oh_groups = (slots + 31) / 32;
        prm.oh_out = oh->out;
        TM_CUDA(cudaMemsetAsync(oh->out, 0, sizeof(float) * (size_t)slots * (size_t)p, st));
    }
    const int half = prm.mtiles * TILE_BYTES;
    prm.f_order = c_order ? 0 : 1;
    prm.tile_mask = 7;
    prm.aux_off = (int)((BK * p * 4 + 127) / 128 * 128);
    if (ps) {
        prm.panel = 1;
        prm.a0 = ps->a0, prm.wa = ps->wa, prm.b0 = ps->b0, prm.wb = ps->wb;
        prm.ldo = (int)p_full;
        prm.tile_mask = ps->mask;
        // panel A box at 0, panel B box one tile (16 KB) further
        prm.aux_off = ps->wb ? TILE_BYTES + (BK * ps->wb * 4 + 127) / 128 * 128
                             : (BK * ps->wa * 4 + 127) / 128 * 128;
    }
    prm.r_bytes = prm.aux_off + 128 + 8 * 128 + 128;   // X tile | d | 8 code vectors | v
    prm.has_v = v ? 1 : 0;
    prm.vec_out = vec_out;
    prm.nsub = g_dense_f32_mode == 3 ? 3 : 1;
    prm.round_mode = tc_round_mode();
    if (prm.f_order) prm.r_bytes = (prm.r_bytes + 1023) / 1024 * 1024;  // swizzle atom = 8 x 128 B
    int scw = 0;
    if (scatter && (scatter->n_cat > 0 || scatter->out_sparse)) {
        if (!dense_tc_scatter_eligible(p, scatter->n_cat, scatter->out_sparse != nullptr))
            return fail("dense_tc: scatter work not eligible for the fused form");
        scw = tc_scatter_warps(scatter->out_sparse != nullptr);
        prm.sc_ncat = scatter->n_cat;
        for (int c = 0; c < scatter->n_cat; ++c) {
            prm.sc_codes[c] = scatter->codes[c];
            prm.sc_tab[c] = static_cast<float*>(scatter->tab[c]);
            prm.sc_K[c] = scatter->K[c];
            prm.sc_copies[c] = scatter->copies[c] > 0 ? scatter->copies[c] : 1;
            prm.sc_df[c] = scatter->drop_first[c];
        }
        // cats only: stage their code vectors by TMA after the one-hot ones when the 8 slots
        // suffice and the vectors are 16-byte aligned
        if (!scatter->out_sparse && prm.oh_ncat + scatter->n_cat <= 8) {
            bool ok = true;
            for (int c = 0; c < scatter->n_cat; ++c)
                ok = ok && (reinterpret_cast<uintptr_t>(scatter->codes[c]) & 15) == 0;
            if (ok) {
                for (int c = 0; c < scatter->n_cat; ++c)
                    if (!encode_1d(&tmaps.codes[prm.oh_ncat + c], scatter->codes[c],
                                   CU_TENSOR_MAP_DATA_TYPE_INT32))
                        return fail("cuTensorMapEncodeTiled failed (scatter codes)");
                prm.sc_staged = 1;
            }
        }
        prm.csr_data = static_cast<const float*>(scatter->csr_data);
        prm.csr_indices = scatter->csr_indices;
        prm.csr_indptr = scatter->csr_indptr;
        prm.out_sparse = static_cast<float*>(scatter->out_sparse);
    }
    // operand-ring depth: as deep as TMEM (512 columns) and shared memory (>= 3 raw stages) allow
    // (TABMAT_B200_TC_SB caps it: a shallower operand ring leaves room for more raw stages)
    static const int sb_cap = [] {
        const char* e = getenv("TABMAT_B200_TC_SB");
        int v = e ? atoi(e) : MAX_SB;
        return v < 2 ? 2 : (v > MAX_SB ? MAX_SB : v);
    }();
    int sb = sb_cap;
    while (sb > 2 && (ntiles * 128 + prm.oh_groups * 32 + sb * prm.mtiles * 32 > 512 ||
                      (SMEM_BUDGET - sb * (half + prm.oh_groups * GROUP_BYTES) - 1536) / prm.r_bytes < 3))
        --sb;
    prm.sb = sb;
    int cols_needed = ntiles * 128 + prm.oh_groups * 32 + sb * prm.mtiles * 32;
    prm.tmem_cols = cols_needed <= 128 ? 128 : (cols_needed <= 256 ? 256 : 512);
    const int fixed = sb * (half + prm.oh_groups * GROUP_BYTES) + 1024 /*align*/ + 512 /*barriers*/;
    int stagesR = (SMEM_BUDGET - fixed) / prm.r_bytes;
    if (stagesR > MAX_STAGES) stagesR = MAX_STAGES;
    if (stagesR < 2) return fail("dense_tc: not enough shared memory for 2 stages");
    prm.stagesR = stagesR;
    size_t smem = (size_t)stagesR * prm.r_bytes + fixed;

    if (!ps) TM_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(p * p), st));
    // TABMAT_B200_TC_NSW = 8 | 16 scale warps
    static const int nsw = getenv("TABMAT_B200_TC_NSW") ? atoi(getenv("TABMAT_B200_TC_NSW")) : 8;
    // The tensor core's fp32 accumulator TRUNCATES at every K = 8 step: a chain of T steps onto a
    // growing sum loses ~T * 2^-25 of it (measured: -9.3e-4 on the diagonal of X^T D X at n = 4e7
    // on one GPU = 34k steps per CTA, against 6.5e-5 for the fp32 CUDA-core kernel).  So a long
    // input is cut into launches of at most `chain` steps per CTA; each launch drains its TMEM
    // accumulators into `out` with fp32 REDs (rounded to nearest).  The kernel's loops stay
    // untouched (cutting the chain inside the kernel cost 0.9-1.7 ms on the 7 ms pass even with
    // the cut disabled, profiles/bench_r3f..r3i_*); an extra launch costs ~0.03 ms
    // (4096 steps: 8 launches at n = 4e7, error 1.2e-4; 2048: 16 launches, 5.8e-5, +0.3 ms).
    const long long total_tiles = prm.num_row_tiles;
    long long seg_tiles = total_tiles;
    {
        static const int env_steps = getenv("TABMAT_B200_TC_FLUSH_STEPS") ? atoi(getenv("TABMAT_B200_TC_FLUSH_STEPS")) : 4096;
        const long long steps = g_tc_flush_steps >= 0 ? g_tc_flush_steps : env_steps;
        if (steps > 0) {
            long long per_cta = steps / ((BK / 8) * prm.nsub);
            if (per_cta < 1) per_cta = 1;
            seg_tiles = per_cta * sm_count();
        }
    }
    int rc = 0;
    for (long long t0 = 0; t0 < total_tiles && rc == 0; t0 += seg_tiles) {
    prm.tile0 = t0;
    prm.num_row_tiles = total_tiles - t0 < seg_tiles ? total_tiles - t0 : seg_tiles;
    // an almost empty last launch is folded into the previous one
    if (total_tiles - t0 - prm.num_row_tiles < seg_tiles / 4) prm.num_row_tiles = total_tiles - t0;
    // SMs left to a collective that runs beside this kernel (tm_set_tc_sm_reserve)
    const long long sms = sm_count() - g_tc_sm_reserve > 8 ? sm_count() - g_tc_sm_reserve : 8;
    long long grid = prm.num_row_tiles < sms ? prm.num_row_tiles : sms;
    if (prm.nsub == 3) {
        if (scw == 8)
            rc = launch_tc<1, 8, 3>(tmaps, prm, (unsigned)grid, smem, st);
        else if (scw == 4)
            rc = launch_tc<1, 4, 3>(tmaps, prm, (unsigned)grid, smem, st);
        else
            rc = launch_tc<1, 0, 3>(tmaps, prm, (unsigned)grid, smem, st);
    } else if (scw == 8) {
        rc = launch_tc<1, 8, 1>(tmaps, prm, (unsigned)grid, smem, st);
    } else if (scw == 4 && nsw == 16) {
        rc = launch_tc<1, 4, 1, 16>(tmaps, prm, (unsigned)grid, smem, st);
    } else if (scw == 4) {
        rc = launch_tc<1, 4, 1>(tmaps, prm, (unsigned)grid, smem, st);
    } else if (nsw == 16 && !share_sm) {
        rc = launch_tc<1, 0, 1, 16>(tmaps, prm, (unsigned)grid, smem, st);
    } else if (share_sm) {
        rc = launch_tc<2, 0, 1>(tmaps, prm, (unsigned)grid, smem, st);
    } else {
        rc = launch_tc<1, 0, 1>(tmaps, prm, (unsigned)grid, smem, st);
    }
    if (prm.num_row_tiles == total_tiles - t0) break;
    }
    if (rc) return rc;
    if (ps) return 0;   // the caller mirrors the triangle once, after the last panel pass
    return symmetrize_from_upper<float>(out, p, st);
}

}  // namespace tmb

extern "C" {

int tm_dense_onehot_sandwich_f32(const float* X, int64_t n, int64_t p, const float* d,
                                 const int32_t* rows, int64_t n_rows, int n_cat,
                                 const int32_t* const* codes, const int64_t* K,
                                 const int32_t* drop_first, float* out_dense, float* out_cat,
                                 tm_stream_t stream) {
    using namespace tmb;
    cudaStream_t st = as_stream(stream);
    if (!dense_tc_eligible(n, p, 1, X) || p > 128)
        return fail("tm_dense_onehot_sandwich_f32: needs sm_100, row-major X, p % 4 == 0, p <= 128");
    if (n_cat < 0 || n_cat > 8) return fail("tm_dense_onehot_sandwich_f32: 0..8 categorical blocks");
    TcOneHot oh;
    oh.ncat = n_cat;
    oh.out = out_cat;
    int64_t slots = 0;
    for (int c = 0; c < n_cat; ++c) {
        oh.codes[c] = codes[c];
        oh.K[c] = (int)K[c];
        oh.drop_first[c] = drop_first[c];
        slots += K[c];
    }
    if (slots > TC_ONEHOT_MAX_SLOTS)
        return fail("tm_dense_onehot_sandwich_f32: more than 320 category columns in total");
    if (rows) {
        Scratch dm(sizeof(float) * (size_t)n, st);
        if (dm.err != cudaSuccess) return fail_cuda(dm.err, "scratch");
        int rc = masked_weights<float>(d, n, rows, n_rows, dm.as<float>(), st);
        if (rc) return rc;
        return dense_sandwich_tc_f32(X, n, p, 1, dm.as<float>(), out_dense, st, n_cat ? &oh : nullptr);
    }
    return dense_sandwich_tc_f32(X, n, p, 1, d, out_dense, st, n_cat ? &oh : nullptr);
}

int tm_has_tcgen05(void) {
    return tmb::tc::device_cc_major() == 10 && tmb::tc::get_encode() != nullptr ? 1 : 0;
}
void tm_set_dense_f32_mode(int mode) { tmb::g_dense_f32_mode = mode; }
void tm_set_tc_round_mode(int mode) { tmb::tc::g_tc_round = mode; }
void tm_set_tc_flush_steps(int steps) { tmb::tc::g_tc_flush_steps = steps; }
void tm_set_tc_sm_reserve(int sms) { tmb::tc::g_tc_sm_reserve = sms < 0 ? 0 : (sms > 64 ? 64 : sms); }
void tm_set_tc_scatter_warps(int warps) {
    if (warps == -1 || warps == 0 || warps == 4 || warps == 8) tmb::g_tc_scatter_warps = warps;
}
/* test hook (not part of the public header): device buffer of >= 65536 + 3*128*128 floats */
void tm_debug_set_tc_buffer(float* buf) { tmb::g_tc_dbg = buf; }
void tm_debug_set_tc_variant(int v) { tmb::g_tc_variant = v; }

}  // extern "C"
