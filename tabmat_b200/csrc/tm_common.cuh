// Shared helpers for the tabmat_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>

#include "../../include/tabmat_b200.h"

namespace tmb {

// ---- error plumbing -------------------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return 1;
}
inline int fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return 2;
}

#define TM_CUDA(call)                                             \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return tmb::fail_cuda(_e, #call);   \
    } while (0)

#define TM_LAUNCHED()                                                     \
    do {                                                                  \
        tmb::g_launches.fetch_add(1, std::memory_order_relaxed);           \
        cudaError_t _e = cudaGetLastError();                              \
        if (_e != cudaSuccess) return tmb::fail_cuda(_e, "kernel launch"); \
    } while (0)

inline cudaStream_t as_stream(tm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// ---- stream-ordered scratch (device workspace arena; replaces reference alloc.h) -------
// The default memory pool gives unused memory back to the driver at every synchronisation
// (release threshold 0), so a caller that synchronises between calls paid a fresh physical
// allocation per scratch buffer (~0.1-0.2 ms each).  Keep it cached instead.
void keep_pool_memory();
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t err = cudaSuccess;
    Scratch(size_t bytes, cudaStream_t st) : s(st) {
        if (bytes == 0) bytes = 16;
        keep_pool_memory();
        err = cudaMallocAsync(&p, bytes, st);
    }
    ~Scratch() {
        if (p) cudaFreeAsync(p, s);
    }
    template <typename T>
    T* as() {
        return reinterpret_cast<T*>(p);
    }
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
};

// ---- device helpers -------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void red_add(F* addr, F v) {
    atomicAdd(addr, v);  // result unused -> RED.E.ADD
}

__device__ __forceinline__ int64_t row_at(const int32_t* __restrict__ rows, int64_t t) {
    return rows ? (int64_t)rows[t] : t;
}

constexpr int kMaxGridX = 2147483647;

inline int grid_for(int64_t work_items, int block, int max_blocks) {
    int64_t g = (work_items + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

// out[i] = 0 for i < n (tiny helper used where cudaMemsetAsync would also do)
template <typename F>
__global__ void k_fill_zero(F* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = F(0);
}

// map[j] = position of j in list (or -1).  map has `p` entries, pre-filled with -1.
__global__ void k_build_pos_map(const int32_t* __restrict__ list, int64_t m,
                                int32_t* __restrict__ map);
// mask[j] = 1 for j in list
__global__ void k_build_mask(const int32_t* __restrict__ list, int64_t m,
                             uint8_t* __restrict__ mask);

// Build an int32 position map of length p on `stream`: identity when list == nullptr is NOT
// materialised (callers test the pointer).  Returns 0 on success.
int build_pos_map(const int32_t* list, int64_t m, int64_t p, int32_t* map, cudaStream_t st);
int build_mask(const int32_t* list, int64_t m, int64_t p, uint8_t* mask, cudaStream_t st);

// dmask[k] = d[k] if k in rows else 0   (row restriction folded into the weights)
template <typename F>
int masked_weights(const F* d, int64_t n, const int32_t* rows, int64_t n_rows, F* dmask,
                   cudaStream_t st);

// mirror the strict lower triangle of a square row-major matrix into the upper one
template <typename F>
int symmetrize_from_lower(F* out, int64_t m, cudaStream_t st);
template <typename F>
int symmetrize_from_upper(F* out, int64_t m, cudaStream_t st);

// ---- tcgen05 dense kernel (dense_tc.cu) -----------------------------------------------
// Categorical blocks with few levels that ride along the fp32 weighted SYRK as one-hot MMAs:
// out[(off_c + codes_c[k] - drop_first_c) * p + b] += d[k] * X[k, b]   (off_c = sum of K before c)
constexpr int TC_ONEHOT_MAX_SLOTS = 320;  // TMEM: 128 (SYRK) + 320 + 64 (operand ring) = 512
struct TcOneHot {
    int ncat;
    const int32_t* codes[8];
    int K[8];
    int drop_first[8];
    float* out;  // [sum K][p], overwritten
};
// Scatter work that can ride along the tcgen05 kernel (dense_tc.cu, "fused" form): the
// dense x many-level categorical blocks (run-aggregated vector REDs) and dense x sparse (one
// vector RED per non-zero) are issued by extra warps from the TMA-staged X tile, so the dense
// block is read from HBM once per sandwich.  Filled by cross_prepare (split_fused.cu).
constexpr int FC_MAX_CATS = 8;
struct FusedCrossParams {
    const int32_t* codes[FC_MAX_CATS];
    void* tab[FC_MAX_CATS];       // K_i * copies_i rows of P values (scratch when copies_i > 1)
    int K[FC_MAX_CATS];
    int copies[FC_MAX_CATS];
    int drop_first[FC_MAX_CATS];
    int n_cat;
    const void* csr_data;
    const int32_t* csr_indices;
    const int32_t* csr_indptr;
    void* out_sparse;             // p_sparse x P, or nullptr
};
constexpr int TC_SCATTER_MAX_CATS = 4;
int cat_dense_gather_f32(const float* X, int64_t p, const float* d, const int32_t* perm,
                         const int32_t* segptr, int64_t K, int64_t n_valid, float* out,
                         cudaStream_t st);
int dense_sandwich_tc_f32(const float* X, int64_t n, int64_t p, int c_order, const float* d,
                          float* out, cudaStream_t st, const TcOneHot* oh = nullptr,
                          bool share_sm = false, const FusedCrossParams* scatter = nullptr,
                          const float* v = nullptr, float* vec_out = nullptr);
// does the tcgen05 kernel take the scatter work of `n_cat` many-level blocks along (p <= 128)?
bool dense_tc_scatter_eligible(int64_t p, int n_cat, bool with_sparse);
bool dense_tc_eligible(int64_t n, int64_t p, int c_order, const void* X);

// ---- fused dense-operand cross blocks (split_fused.cu) ---------------------------------
// `runs` != 0 selects the run-aggregating kernel (consecutive rows per warp, one RED per run of
// equal codes; needs <= 4 categorical blocks).  g_cross_runs_mode: 0 = as asked, 1 = always,
// 2 = never.
template <typename F>
int dense_cross_fused(const F* X, int64_t n, int64_t p, const F* d, const int32_t* rows,
                      int64_t n_rows, int n_cat, const int32_t* const* codes, const int64_t* K,
                      const int32_t* drop_first, F* const* out_cat, const F* csr_data,
                      const int32_t* csr_indices, const int32_t* csr_indptr, int64_t p_sparse,
                      F* out_sparse, int runs, cudaStream_t st);
extern int g_cross_runs_mode;
extern int g_sm_reserve;   // SMs the gather kernel leaves free (tm_set_sm_reserve)
// dense x sparse by row-blocked gather (split_fused.cu): out (p_s x p) overwritten
template <typename F>
int csc_dense_gather(const F* X, int64_t p, const F* d, const F* bdata, const int32_t* brow,
                     const int32_t* bptr, int64_t p_s, int64_t n_blocks, F* out, cudaStream_t st);
// The two halves of dense_cross_fused around the kernel launch: zero-fill the destinations,
// set up the replicated tables of few-level blocks (scratch handed back through `scr`, the
// caller owns it until cross_finish has been enqueued), and sum the replicas afterwards.
struct CrossScratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~CrossScratch() {
        if (p) cudaFreeAsync(p, s);
    }
};
template <typename F>
int cross_prepare(int64_t p, int n_cat, const int32_t* const* codes, const int64_t* K,
                  const int32_t* drop_first, F* const* out_cat, const F* csr_data,
                  const int32_t* csr_indices, const int32_t* csr_indptr, int64_t p_sparse,
                  F* out_sparse, FusedCrossParams& prm, CrossScratch& scr, cudaStream_t st);
template <typename F>
int cross_finish(int64_t p, int n_cat, const int64_t* K, F* const* out_cat,
                 const FusedCrossParams& prm, cudaStream_t st);
constexpr int TM_BLOCK_FLAG_RUNS = 1;     // tm_block_desc.flags bit 0
constexpr int TM_BLOCK_FLAG_PRIMARY = 2;  // bit 1: the primary sort key
extern int g_dense_f32_mode;

// ---- fused index blocks (split_index.cu): categorical self / pair blocks and categorical x
// sparse from 32-byte row records {d, codes}; see the header comment there ------------------
template <typename F>
bool index_fused_eligible(int n_cat, const int64_t* K);
template <typename F>
size_t index_record_bytes(int64_t n);
template <typename F>
int index_pack_records(const F* d, int64_t n, int n_cat, const int32_t* const* codes,
                       const int32_t* drop_first, void* rec, cudaStream_t st);
// rec == NULL: d and the code vectors are read directly (no packing pass)
template <typename F>
int index_cat_pairs(const void* rec, const F* d, const int32_t* const* codes,
                    const int32_t* drop_first, int64_t n, int n_cat, const int64_t* K,
                    const int32_t* runs, F* const* outs_self, F* const* outs_pair,
                    cudaStream_t st);
template <typename F>
bool index_cat_sparse_fits(int n_cat, const int64_t* K, int64_t p_s);
// bit-packed codes of the CSC non-zeros' rows (tm_block_desc.csc_cat_codes): field widths
int index_pack_width(int64_t ncols);
bool index_pack_fits(int n_cat, const int64_t* K);
// rec == NULL: `packed` + `d` replace the row records
template <typename F>
int index_cat_sparse(const void* rec, const F* d, const uint64_t* packed, int n_cat,
                     const int64_t* K, const int32_t* runs, const F* csc_data,
                     const int32_t* csc_row, const int32_t* csc_indptr, int64_t p_s,
                     int n_row_blocks, F* const* outs, F* diag, int64_t diag_ld, cudaStream_t st);
// sparse self sandwich (sparse.cu); offdiag_only: `out` was zero-filled by the caller and its
// diagonal comes from elsewhere (index_cat_sparse's by-product), the kernel adds the strictly
// lower triangle and mirrors it
template <typename F>
int sparse_sandwich_ex(const F* data, const int32_t* indices, const int32_t* indptr,
                       const int32_t* nz_row, int64_t n, int64_t p, int64_t nnz, const F* d,
                       const int32_t* rows, int64_t n_rows, const int32_t* cols, int64_t m, F* out,
                       cudaStream_t st, bool offdiag_only);

}  // namespace tmb
