// Fused "index" blocks of a SplitMatrix sandwich: everything that involves only categorical
// codes, the weights d and the sparse block.
//
// The reference computes these one pair at a time from Python (split_matrix.py:346-354):
//   categorical self     categorical.pyx:183-218   diag[c_i[k]] += d[k]
//   categorical_i x _j   split.pyx:83-111          out[c_i[k], c_j[k]] += d[k]
//   categorical x sparse categorical_matrix.py:825-838 (scipy csr_matmat of the transposed
//                        one-hot matrix with the CSC block)   out[c_i[k], j] += d[k] * A[k, j]
// The first-generation kernels here (categorical.cu) did the same pair by pair, each re-reading
// its code vectors and d, each sending one L2 RED per row or per non-zero: 20 launches and
// ~1e9 scalar REDs per step at the benchmark shape, the longest of the three passes.
//
// This file does it in three launches:
//   k_pack_records     rec[k] = { d[k] (0 outside `rows`), c_0[k]-drop_first_0, c_1[k]-..., ... }
//                      one 32-byte record per row = one L2 sector per gather;
//   k_cat_pairs        one pass over the records: every categorical self block and every
//                      categorical x categorical block; tables that fit go to shared memory
//                      (tiny ones replicated per lane), the rest are L2 REDs with runs of equal
//                      keys pre-summed in registers (row-sorted storage, row_order.py);
//   k_cat_sparse_csc   categorical x sparse for ALL categorical blocks, driven by the CSC copy
//                      of the sparse block: a CTA owns a column j, walks its non-zeros, gathers
//                      the 32-byte record of each row and adds d[k] * A[k, j] to out_i[c_i, j]:
//                      per-thread private shared-memory tables for blocks with few levels,
//                      shared-memory atomics for mid-sized ones, L2 REDs for the widest
//                      (8 bytes per non-zero streamed + one sector per non-zero gathered);
//   k_cat_sparse_cols  the same over a ROW-BLOCKED CSC copy (opt-in): a CTA owns a few columns
//                      for the whole kernel and all CTAs sweep the row blocks at the same pace,
//                      so the record gathers stay inside L2.
#include <cstdlib>

#include "tm_common.cuh"

namespace tmb {

constexpr int IDX_MAX_CATS_DECL = 7;
template <typename F>
struct RowRec;
template <>
struct __align__(32) RowRec<float> {
    float d;
    int32_t c[7];
};
template <>
struct __align__(32) RowRec<double> {
    double d;
    int32_t c[6];
};
template <typename F>
constexpr int rec_max_cats() {
    return (32 - (int)sizeof(F)) / 4;
}

struct PackParams {
    const int32_t* codes[7];
    int drop_first[7];
};

template <typename F, int NC>
__global__ void k_pack_records(const F* __restrict__ d, int64_t n, const PackParams prm,
                               RowRec<F>* __restrict__ rec) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        RowRec<F> r;
        r.d = d[k];
#pragma unroll
        for (int c = 0; c < rec_max_cats<F>(); ++c) r.c[c] = -1;
        // every term of row k is proportional to d[k]: rows with d[k] == 0 (and the rows
        // outside a `rows` restriction, whose weight the caller zeroed) are marked missing
        if (r.d != F(0)) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int v = prm.codes[c][k] - prm.drop_first[c];
                r.c[c] = v < 0 ? -1 : v;
            }
        }
        // two 16-byte stores
        const int4* src = reinterpret_cast<const int4*>(&r);
        int4* dst = reinterpret_cast<int4*>(rec + k);
        dst[0] = src[0];
        dst[1] = src[1];
    }
}

template <typename F>
__device__ __forceinline__ RowRec<F> load_rec(const RowRec<F>* __restrict__ rec, int64_t k) {
    RowRec<F> r;
    const int4* src = reinterpret_cast<const int4*>(rec + k);
    int4* dst = reinterpret_cast<int4*>(&r);
    dst[0] = __ldg(src);
    dst[1] = __ldg(src + 1);
    return r;
}

// Where a CSC-driven kernel takes the weight and the categorical codes of a non-zero's row from:
//   records   rec[k] = {d, codes}: one 32-byte gather per non-zero (k_pack_records per call);
//   packed    pk[e] = the codes of non-zero e's row, bit-packed into 64 bits in CSC order (built
//             ONCE per matrix, like the cached CSR: streamed, coalesced) + a 4/8-byte gather of
//             d[k].  No per-call packing pass, a quarter of the gathered bytes, and with the
//             row-blocked CSC order the gather window (block_rows * sizeof F) sits in L2.
struct CodeSrc {
    const void* rec;
    const void* d;
    const unsigned long long* pk;
    int shift[IDX_MAX_CATS_DECL];
    unsigned mask[IDX_MAX_CATS_DECL];   // (1 << width) - 1 = the "missing" marker
};

template <typename F, int NC, bool PK>
__device__ __forceinline__ void fetch_row(const CodeSrc& src, int e, int k, F& dk, int (&key)[NC]) {
    if (PK) {
        dk = __ldg(static_cast<const F*>(src.d) + k);
        const unsigned long long pk = __ldg(src.pk + e);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const unsigned v = (unsigned)(pk >> src.shift[c]) & src.mask[c];
            key[c] = (v == src.mask[c] || dk == F(0)) ? -1 : (int)v;
        }
    } else {
        const RowRec<F> r = load_rec<F>(static_cast<const RowRec<F>*>(src.rec), k);
        dk = r.d;
#pragma unroll
        for (int c = 0; c < NC; ++c) key[c] = r.c[c];
    }
}

// Sum `val` over maximal runs of consecutive lanes with equal keys; true on the first lane of a
// run, whose val then holds the run total (same helper as in categorical.cu).
template <typename F, typename KeyT>
__device__ __forceinline__ bool run_reduce(KeyT key, F& val, int lane) {
    const unsigned FULL = 0xffffffffu;
    const KeyT prev = __shfl_up_sync(FULL, key, 1);
    const bool head = lane == 0 || prev != key;
    const unsigned heads = __ballot_sync(FULL, head);
    if (heads != FULL) {
        const unsigned above = heads & ~((2u << lane) - 1u);
        const int end = above ? __ffs(above) - 2 : 31;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const F o = __shfl_down_sync(FULL, val, off);
            if (lane + off <= end) val += o;
        }
    }
    return head;
}

// ---------------------------------------------------------------------------------------
// categorical self + categorical x categorical, one pass over the records
// ---------------------------------------------------------------------------------------
constexpr int IDX_MAX_CATS = 7;
constexpr int IDX_MAX_TARGETS = IDX_MAX_CATS + IDX_MAX_CATS * (IDX_MAX_CATS - 1) / 2;  // 28

struct PairParams {
    int K[IDX_MAX_CATS];
    int runs[IDX_MAX_CATS];          // rows are stored sorted by this block's codes
    // target t: t < NC the self block of cat t; then the pairs (i, j), i < j, in row-major order
    void* out[IDX_MAX_TARGETS];      // global destination (zero-filled by the host)
    int smem_off[IDX_MAX_TARGETS];   // element offset of the shared-memory table, or -1
    int copies[IDX_MAX_TARGETS];     // replicas of the shared-memory table (lane % copies)
    int smem_elems;                  // total shared-memory elements
};

constexpr int PAIRS_THREADS = 1024;

template <typename F>
__device__ __forceinline__ void pair_add(const PairParams& prm, F* smem, int t, long long key,
                                         long long size, F val, bool runs, int lane) {
    bool head = true;
    if (runs) head = run_reduce<F, long long>(key, val, lane);
    if (!head || key < 0) return;
    if (prm.smem_off[t] >= 0) {
        if (prm.copies[t] == PAIRS_THREADS)  // one replica per thread, [key][thread]: plain RMW
            smem[prm.smem_off[t] + key * PAIRS_THREADS + threadIdx.x] += val;
        else
            atomicAdd(smem + prm.smem_off[t] + (long long)(lane % prm.copies[t]) * size + key,
                      val);
    } else {
        red_add(static_cast<F*>(prm.out[t]) + key, val);
    }
}

// DIRECT: the codes and d are read from their own arrays (coalesced, 4 (NC + 1) bytes per row)
// instead of from packed 32-byte records
template <typename F, int NC, int U, bool DIRECT>
__global__ void __launch_bounds__(PAIRS_THREADS, 1)
k_cat_pairs(const RowRec<F>* __restrict__ rec, const F* __restrict__ dvec, const PackParams src,
            int64_t n, const PairParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* smem = reinterpret_cast<F*>(smem_raw);
    for (int i = threadIdx.x; i < prm.smem_elems; i += blockDim.x) smem[i] = F(0);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = gw * (32 * U); base < n; base += nw * (32 * U)) {
        RowRec<F> r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t k = base + u * 32 + lane;
            if (k < n && DIRECT) {
                r[u].d = dvec[k];
#pragma unroll
                for (int c = 0; c < rec_max_cats<F>(); ++c) r[u].c[c] = -1;
                if (r[u].d != F(0)) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int v = __ldg(src.codes[c] + k) - src.drop_first[c];
                        r[u].c[c] = v < 0 ? -1 : v;
                    }
                }
            } else if (k < n) {
                r[u] = load_rec<F>(rec, k);
            } else {
                r[u].d = F(0);
#pragma unroll
                for (int c = 0; c < rec_max_cats<F>(); ++c) r[u].c[c] = -1;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const F dk = r[u].d;
            int t = NC;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int ci = r[u].c[i];
                pair_add<F>(prm, smem, i, (long long)ci, (long long)prm.K[i], dk,
                            prm.runs[i] != 0, lane);
#pragma unroll
                for (int j = i + 1; j < NC; ++j) {
                    const int cj = r[u].c[j];
                    const long long key =
                        (ci >= 0 && cj >= 0) ? (long long)ci * prm.K[j] + cj : -1ll;
                    pair_add<F>(prm, smem, t, key, (long long)prm.K[i] * prm.K[j], dk,
                                prm.runs[i] != 0 && prm.runs[j] != 0, lane);
                    ++t;
                }
            }
        }
    }
    __syncthreads();
    // flush the shared-memory tables (replicas summed) with REDs
    constexpr int NT = NC + NC * (NC - 1) / 2;
    int t = NC;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
#pragma unroll
        for (int jj = i; jj < NC; ++jj) {  // jj == i: the self block (target i)
            const int tt = jj == i ? i : t++;
            if (tt >= NT || prm.smem_off[tt] < 0) continue;
            const int size = jj == i ? prm.K[i] : prm.K[i] * prm.K[jj];  // fits: it is in smem
            const F* tab = smem + prm.smem_off[tt];
            F* out = static_cast<F*>(prm.out[tt]);
            if (prm.copies[tt] == PAIRS_THREADS) {  // per-thread replicas: a warp sums one entry
                for (int e = threadIdx.x >> 5; e < size; e += PAIRS_THREADS / 32) {
                    F sum = F(0);
                    for (int q = lane; q < PAIRS_THREADS; q += 32) sum += tab[e * PAIRS_THREADS + q];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    if (lane == 0 && sum != F(0)) red_add(out + e, sum);
                }
                continue;
            }
            for (int e = threadIdx.x; e < size; e += blockDim.x) {
                F s = F(0);
                for (int c = 0; c < prm.copies[tt]; ++c) s += tab[(long long)c * size + e];
                if (s != F(0)) red_add(out + e, s);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// categorical x sparse for all categorical blocks, CSC-driven, shared-memory column tables
// ---------------------------------------------------------------------------------------
enum { CS_PRIV = 0, CS_ATOM = 1, CS_L2 = 2 };
constexpr int CS_THREADS = 256;

struct CatSparseParams {
    int K[IDX_MAX_CATS];
    int mode[IDX_MAX_CATS];      // CS_PRIV: one replica per thread, [level][thread], plain RMW
                                 // CS_ATOM: shared-memory atomics, `rep` replicas (lane % rep)
                                 // CS_L2:   scalar REDs straight into out (zero-filled by the host)
    int rep[IDX_MAX_CATS];
    int off[IDX_MAX_CATS];       // element offset of block c's shared-memory table
    int runs[IDX_MAX_CATS];
    void* out[IDX_MAX_CATS];     // K_c x p_s, row-major
    int smem_elems;
};

// One CTA per sparse column j (grid-stride), plain CSC.  Per non-zero (k, j, a): val = d[k] * a
// from the 32-byte row record, then per categorical block one update of out_c[code_c[k], j]:
//   few levels   -> the thread's private column table in shared memory (no atomics),
//   some levels  -> shared-memory atomics on a replicated column table,
//   many levels  -> scalar L2 RED (the addresses of one column are p_s * 4 bytes apart, so they
//                   never share a sector; runs of equal codes are pre-summed when the rows are
//                   stored sorted with that block as the primary key).
// ncu (profiles/ncu_step_r1c_summary.csv): 3.6 ms at the benchmark shape, latency-bound on the
// record gathers — 8.7 GB of DRAM reads, i.e. one 64-byte access per non-zero; the row-blocked
// variant below keeps them in L2.
template <typename F, int NC, int U, bool PK>
__global__ void __launch_bounds__(CS_THREADS)
k_cat_sparse_csc(const F* __restrict__ data, const int32_t* __restrict__ row_idx,
                 const int32_t* __restrict__ indptr, int p_s, const CodeSrc src,
                 const CatSparseParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* smem = reinterpret_cast<F*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    constexpr int NW = CS_THREADS / 32;
    for (int j = blockIdx.x; j < p_s; j += gridDim.x) {
        for (int i = threadIdx.x; i < prm.smem_elems; i += CS_THREADS) smem[i] = F(0);
        __syncthreads();
        const int e0 = indptr[j], e1 = indptr[j + 1];
        for (int eb = e0 + wib * (32 * U); eb < e1; eb += NW * (32 * U)) {
            F dk[U];
            int keys[U][NC];
            F a[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = eb + u * 32 + lane;
                a[u] = F(0);
                int k = -1;
                if (e < e1) {
                    k = row_idx[e];
                    a[u] = data[e];
                }
                if (k >= 0) {
                    fetch_row<F, NC, PK>(src, e, k, dk[u], keys[u]);
                } else {
                    dk[u] = F(0);
#pragma unroll
                    for (int c = 0; c < NC; ++c) keys[u][c] = -1;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const F val0 = dk[u] * a[u];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    int key = keys[u][c];
                    if (prm.mode[c] == CS_PRIV) {
                        if (key >= 0) smem[prm.off[c] + key * CS_THREADS + threadIdx.x] += val0;
                        continue;
                    }
                    F val = val0;
                    bool head = true;
                    if (prm.runs[c]) head = run_reduce<F, int>(key, val, lane);
                    if (!head || key < 0) continue;
                    if (prm.mode[c] == CS_ATOM)
                        atomicAdd(smem + prm.off[c] + (lane % prm.rep[c]) * prm.K[c] + key, val);
                    else
                        red_add(static_cast<F*>(prm.out[c]) + (int64_t)key * p_s + j, val);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            F* out = static_cast<F*>(prm.out[c]);
            const int Kc = prm.K[c];
            if (prm.mode[c] == CS_PRIV) {
                for (int lvl = wib; lvl < Kc; lvl += NW) {
                    const F* tab = smem + prm.off[c] + lvl * CS_THREADS;
                    F sum = F(0);
#pragma unroll
                    for (int q = 0; q < NW; ++q) sum += tab[q * 32 + lane];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    if (lane == 0) out[(int64_t)lvl * p_s + j] = sum;
                }
            } else if (prm.mode[c] == CS_ATOM) {
                const int rep = prm.rep[c];
                for (int lvl = threadIdx.x; lvl < Kc; lvl += CS_THREADS) {
                    const F* tab = smem + prm.off[c] + lvl;
                    F sum = F(0);
                    for (int q = 0; q < rep; ++q) sum += tab[q * Kc];
                    out[(int64_t)lvl * p_s + j] = sum;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// categorical x sparse over the ROW-BLOCKED CSC copy, column-owner form.
// A CTA owns a contiguous set of sparse columns for the whole kernel and keeps their column
// tables (blocks with <= 512 levels, replicated) in shared memory; it walks the row blocks in
// order and, inside a row block, the non-zeros of its columns.  The grid is one resident wave,
// every CTA has the same expected work per row block, so all CTAs gather row records from the
// same window of `block_rows * 32` bytes at any time — the gathers hit L2 instead of costing a
// 64-byte DRAM access per non-zero (8.7 GB of DRAM reads per launch in k_cat_sparse_csc above).
// No barriers in the main loop, no zero-fill or flush per work item; wide blocks take L2 REDs.
// ---------------------------------------------------------------------------------------
constexpr int CO_THREADS = 256;

struct ColOwnerParams {
    int K[IDX_MAX_CATS];
    int in_smem[IDX_MAX_CATS];
    int rep[IDX_MAX_CATS];
    int off[IDX_MAX_CATS];       // offset of block c's table inside one column slot
    int runs[IDX_MAX_CATS];
    void* out[IDX_MAX_CATS];
    int slot;                    // shared-memory elements per owned column
    int cols_per_cta;
    int n_row_blocks;
    // optional by-product: the diagonal of the sparse block's own sandwich,
    // diag[j * diag_ld] = sum_k d[k] * A[k, j]^2 - this kernel walks every non-zero of column j
    // with d[k] in hand anyway, and the CSR outer-product kernel then skips its 40 % of REDs
    // that all land on those p_s addresses
    void* diag;
    long long diag_ld;
};

template <typename F, int NC, bool PK>
__global__ void __launch_bounds__(CO_THREADS)
k_cat_sparse_cols(const F* __restrict__ data, const int32_t* __restrict__ row_idx,
                  const int32_t* __restrict__ indptr, int p_s, const CodeSrc src,
                  const ColOwnerParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* smem = reinterpret_cast<F*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * prm.cols_per_cta;
    const int c1 = min(p_s, c0 + prm.cols_per_cta);
    if (c0 >= c1) return;
    F* dsm = smem + (size_t)prm.cols_per_cta * prm.slot;   // [cols_per_cta] diagonal sums
    for (int i = threadIdx.x; i < (c1 - c0) * prm.slot; i += CO_THREADS) smem[i] = F(0);
    for (int i = threadIdx.x; i < prm.cols_per_cta; i += CO_THREADS) dsm[i] = F(0);
    __syncthreads();
    for (int blk = 0; blk < prm.n_row_blocks; ++blk) {
        const int64_t base = (int64_t)blk * p_s;
        for (int j = c0; j < c1; ++j) {
            F* tab = smem + (j - c0) * prm.slot;
            F dacc = F(0);
            const int e0 = indptr[base + j], e1 = indptr[base + j + 1];
            // warp-uniform trip count (run_reduce needs whole warps)
            // two groups of CO_THREADS non-zeros per visit: the loads (and the dependent d
            // gathers) of both are in flight before the first table update
            constexpr int UC = 2;
            for (int eb = e0 + (threadIdx.x & ~31); eb < e1; eb += UC * CO_THREADS) {
                F dk[UC], a[UC];
                int keys[UC][NC];
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    const int e = eb + u * CO_THREADS + lane;
                    dk[u] = F(0);
                    a[u] = F(0);
                    if (e < e1) {
                        a[u] = data[e];
                        fetch_row<F, NC, PK>(src, e, row_idx[e], dk[u], keys[u]);
                    } else {
#pragma unroll
                        for (int c = 0; c < NC; ++c) keys[u][c] = -1;
                    }
                }
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    if (eb + u * CO_THREADS >= e1) break;   // warp-uniform
                    const F val0 = dk[u] * a[u];
                    dacc += val0 * a[u];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        int key = keys[u][c];
                        F val = val0;
                        bool head = true;
                        if (prm.runs[c]) head = run_reduce<F, int>(key, val, lane);
                        if (!head || key < 0) continue;
                        if (prm.in_smem[c])
                            atomicAdd(tab + prm.off[c] + (lane % prm.rep[c]) * prm.K[c] + key, val);
                        else
                            red_add(static_cast<F*>(prm.out[c]) + (int64_t)key * p_s + j, val);
                    }
                }
            }
            if (prm.diag) {   // warp-uniform
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
                if (lane == 0 && dacc != F(0)) atomicAdd(dsm + (j - c0), dacc);
            }
        }
    }
    __syncthreads();
    if (prm.diag)   // this CTA is the only owner of these columns
        for (int i = threadIdx.x; i < c1 - c0; i += CO_THREADS)
            static_cast<F*>(prm.diag)[(int64_t)(c0 + i) * prm.diag_ld] = dsm[i];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (!prm.in_smem[c]) continue;
        F* out = static_cast<F*>(prm.out[c]);
        const int Kc = prm.K[c], rep = prm.rep[c], nco = c1 - c0;
        for (int i = threadIdx.x; i < Kc * nco; i += CO_THREADS) {
            const int lvl = i / nco, jj = i - lvl * nco;  // consecutive threads, consecutive j
            const F* t = smem + jj * prm.slot + prm.off[c] + lvl;
            F sum = F(0);
            for (int q = 0; q < rep; ++q) sum += t[q * Kc];
            out[(int64_t)lvl * p_s + c0 + jj] = sum;
        }
    }
}

// ---- host side ---------------------------------------------------------------------------
static bool index_fused_off() {
    static const bool off =
        getenv("TABMAT_B200_INDEX_FUSED") && atoi(getenv("TABMAT_B200_INDEX_FUSED")) == 0;
    return off;
}

template <typename F>
bool index_fused_eligible(int n_cat, const int64_t* K) {
    if (index_fused_off() || n_cat < 1 || n_cat > rec_max_cats<F>()) return false;
    for (int c = 0; c < n_cat; ++c)
        if (K[c] <= 0 || K[c] > (1 << 24)) return false;
    return true;
}
template bool index_fused_eligible<float>(int, const int64_t*);
template bool index_fused_eligible<double>(int, const int64_t*);

template <typename F>
size_t index_record_bytes(int64_t n) {
    return sizeof(RowRec<F>) * (size_t)(n > 0 ? n : 1);
}
template size_t index_record_bytes<float>(int64_t);
template size_t index_record_bytes<double>(int64_t);

template <typename F, int NC>
static int launch_pack(const F* d, int64_t n, const PackParams& pp, RowRec<F>* rec,
                       cudaStream_t st) {
    k_pack_records<F, NC><<<grid_for(n, 256 * 2, sm_count() * 16), 256, 0, st>>>(d, n, pp, rec);
    return 0;
}

template <typename F>
int index_pack_records(const F* d, int64_t n, int n_cat, const int32_t* const* codes,
                       const int32_t* drop_first, void* rec_v, cudaStream_t st) {
    if (n <= 0) return 0;
    PackParams pp;
    memset(&pp, 0, sizeof(pp));
    for (int c = 0; c < n_cat; ++c) {
        pp.codes[c] = codes[c];
        pp.drop_first[c] = drop_first[c];
    }
    RowRec<F>* rec = static_cast<RowRec<F>*>(rec_v);
    switch (n_cat) {
        case 1: launch_pack<F, 1>(d, n, pp, rec, st); break;
        case 2: launch_pack<F, 2>(d, n, pp, rec, st); break;
        case 3: launch_pack<F, 3>(d, n, pp, rec, st); break;
        case 4: launch_pack<F, 4>(d, n, pp, rec, st); break;
        case 5: launch_pack<F, 5>(d, n, pp, rec, st); break;
        case 6: launch_pack<F, 6>(d, n, pp, rec, st); break;
        default:
            if (rec_max_cats<F>() >= 7) launch_pack<F, rec_max_cats<F>()>(d, n, pp, rec, st);
            break;
    }
    TM_LAUNCHED();
    return 0;
}
template int index_pack_records<float>(const float*, int64_t, int, const int32_t* const*,
                                       const int32_t*, void*, cudaStream_t);
template int index_pack_records<double>(const double*, int64_t, int, const int32_t* const*,
                                        const int32_t*, void*, cudaStream_t);

constexpr size_t IDX_SMEM_BUDGET = 200 * 1024;

template <typename F, int NC>
static int launch_pairs(const RowRec<F>* rec, const F* d, const PackParams& src, int64_t n,
                        const PairParams& prm, cudaStream_t st) {
    constexpr int U = 2;
    const size_t smem = sizeof(F) * (size_t)prm.smem_elems;
    const int g = grid_for(n, PAIRS_THREADS * U, sm_count());
    // per device and cheap: set on every launch rather than cached in a process-wide static
    if (rec) {
        TM_CUDA(cudaFuncSetAttribute(k_cat_pairs<F, NC, U, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)IDX_SMEM_BUDGET));
        k_cat_pairs<F, NC, U, false><<<g, PAIRS_THREADS, smem, st>>>(rec, d, src, n, prm);
    } else {
        TM_CUDA(cudaFuncSetAttribute(k_cat_pairs<F, NC, U, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)IDX_SMEM_BUDGET));
        k_cat_pairs<F, NC, U, true><<<g, PAIRS_THREADS, smem, st>>>(rec, d, src, n, prm);
    }
    TM_LAUNCHED();
    return 0;
}

// outs_self[c]: K_c values; outs_pair[i * n_cat + j] (i < j): K_i x K_j row-major.  Overwrites.
// rec_v != NULL: read the 32-byte row records; NULL: read d and the code vectors directly.
template <typename F>
int index_cat_pairs(const void* rec_v, const F* d, const int32_t* const* codes,
                    const int32_t* drop_first, int64_t n, int n_cat, const int64_t* K,
                    const int32_t* runs, F* const* outs_self, F* const* outs_pair,
                    cudaStream_t st) {
    const RowRec<F>* rec = static_cast<const RowRec<F>*>(rec_v);
    PackParams src;
    memset(&src, 0, sizeof(src));
    for (int c = 0; c < n_cat && c < 7; ++c) {
        src.codes[c] = codes[c];
        src.drop_first[c] = drop_first[c];
    }
    PairParams prm;
    memset(&prm, 0, sizeof(prm));
    int64_t size[IDX_MAX_TARGETS];
    int nt = n_cat;
    for (int c = 0; c < n_cat; ++c) {
        prm.K[c] = (int)K[c];
        prm.runs[c] = runs ? runs[c] : 0;
        prm.out[c] = outs_self[c];
        size[c] = K[c];
    }
    for (int i = 0; i < n_cat; ++i)
        for (int j = i + 1; j < n_cat; ++j) {
            prm.out[nt] = outs_pair[i * n_cat + j];
            size[nt] = K[i] * K[j];
            ++nt;
        }
    for (int t = 0; t < nt; ++t) {
        prm.smem_off[t] = -1;
        prm.copies[t] = 1;
        TM_CUDA(cudaMemsetAsync(prm.out[t], 0, sizeof(F) * (size_t)size[t], st));
    }
    if (n <= 0) return 0;
    // Placement (rates measured with tools/micro/atom_bench.cu on B200): shared-memory float
    // atomics — compare-and-swap loops — sustain 2-3 lane-adds per clock and SM on tables of
    // >= 500 entries (6-9e11 adds/s in aggregate; 0.2 on 10 entries, hence the replicas), scalar
    // L2 REDs 1.9e11/s on tables of >= 200k entries and collapse below ~50k (5.6e10/s on 10k,
    // 7e9/s on 500).  So: smallest tables first into shared memory until it is full, only the
    // rest as L2 REDs; a self block with <= 16 levels gets one replica per thread and plain
    // read-modify-write (6 lane-adds per clock and SM).
    const size_t budget = IDX_SMEM_BUDGET / sizeof(F);
    size_t used = 0;
    bool placed[IDX_MAX_TARGETS] = {false};
    // first pass: which tables fit at all (one replica each), smallest first
    int order[IDX_MAX_TARGETS];
    int n_fit = 0;
    size_t need = 0;
    for (;;) {
        int best = -1;
        for (int t = 0; t < nt; ++t)
            if (!placed[t] && (best < 0 || size[t] < size[best])) best = t;
        if (best < 0) break;
        placed[best] = true;
        if (need + (size_t)size[best] > budget) continue;
        need += (size_t)size[best];
        order[n_fit++] = best;
    }
    // second pass: what is left goes to replicas of the tiny tables
    size_t spare = budget - need;
    for (int q = 0; q < n_fit; ++q) {
        const int t = order[q];
        const size_t sz = (size_t)size[t];
        int copies = 1;
        if (t < n_cat && sz <= 16 && sz * (PAIRS_THREADS - 1) <= spare / 2) {
            copies = PAIRS_THREADS;  // private per thread
        } else {
            while (copies < 32 && sz * copies * 4 <= 2048 && sz * (2 * copies - 1) <= spare)
                copies *= 2;
        }
        spare -= sz * (copies - 1);
        prm.smem_off[t] = (int)used;
        prm.copies[t] = copies;
        used += sz * copies;
    }
    prm.smem_elems = (int)used;
    switch (n_cat) {
        case 1: return launch_pairs<F, 1>(rec, d, src, n, prm, st);
        case 2: return launch_pairs<F, 2>(rec, d, src, n, prm, st);
        case 3: return launch_pairs<F, 3>(rec, d, src, n, prm, st);
        case 4: return launch_pairs<F, 4>(rec, d, src, n, prm, st);
        case 5: return launch_pairs<F, 5>(rec, d, src, n, prm, st);
        case 6: return launch_pairs<F, 6>(rec, d, src, n, prm, st);
        default: return launch_pairs<F, rec_max_cats<F>()>(rec, d, src, n, prm, st);
    }
}
template int index_cat_pairs<float>(const void*, const float*, const int32_t* const*,
                                    const int32_t*, int64_t, int, const int64_t*, const int32_t*,
                                    float* const*, float* const*, cudaStream_t);
template int index_cat_pairs<double>(const void*, const double*, const int32_t* const*,
                                     const int32_t*, int64_t, int, const int64_t*, const int32_t*,
                                     double* const*, double* const*, cudaStream_t);

constexpr size_t CS_SMEM_BUDGET = 72 * 1024;  // three CTAs of 256 threads per SM

template <typename F, int NC>
static int launch_cat_sparse(const F* data, const int32_t* row_idx, const int32_t* indptr,
                             int p_s, const CodeSrc& src, const CatSparseParams& prm,
                             cudaStream_t st) {
    constexpr int U = 2;
    const size_t smem = sizeof(F) * (size_t)(prm.smem_elems > 0 ? prm.smem_elems : 1);
    const int g = p_s < sm_count() * 12 ? p_s : sm_count() * 12;
    if (src.pk) {
        TM_CUDA(cudaFuncSetAttribute(k_cat_sparse_csc<F, NC, U, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)CS_SMEM_BUDGET));
        k_cat_sparse_csc<F, NC, U, true><<<g, CS_THREADS, smem, st>>>(data, row_idx, indptr, p_s,
                                                                      src, prm);
    } else {
        TM_CUDA(cudaFuncSetAttribute(k_cat_sparse_csc<F, NC, U, false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)CS_SMEM_BUDGET));
        k_cat_sparse_csc<F, NC, U, false><<<g, CS_THREADS, smem, st>>>(data, row_idx, indptr, p_s,
                                                                       src, prm);
    }
    TM_LAUNCHED();
    return 0;
}

// Private per-thread tables for <= 64 levels while they fit, shared-memory atomics up to 512
// levels, L2 REDs beyond (and for whatever does not fit).
template <typename F>
static void cat_sparse_layout(int n_cat, const int64_t* K, const int32_t* runs,
                              CatSparseParams& prm) {
    memset(&prm, 0, sizeof(prm));
    const int64_t budget = (int64_t)(CS_SMEM_BUDGET / sizeof(F));
    int64_t used = 0;
    for (int c = 0; c < n_cat; ++c) {
        prm.K[c] = (int)K[c];
        prm.runs[c] = runs ? runs[c] : 0;
        prm.rep[c] = 1;
        prm.mode[c] = CS_L2;
    }
    // smallest blocks first
    bool done[IDX_MAX_CATS] = {false};
    for (int it = 0; it < n_cat; ++it) {
        int best = -1;
        for (int c = 0; c < n_cat; ++c)
            if (!done[c] && (best < 0 || K[c] < K[best])) best = c;
        done[best] = true;
        const int64_t Kb = K[best];
        if (Kb <= 64 && used + Kb * CS_THREADS <= budget - 2048) {
            prm.mode[best] = CS_PRIV;
            prm.off[best] = (int)used;
            used += Kb * CS_THREADS;
        } else if (Kb <= 512) {
            int rep = 1;
            while (rep < 8 && Kb * rep * 2 <= 2048) rep *= 2;
            if (used + Kb * rep > budget) rep = 1;
            if (used + Kb * rep > budget) continue;  // stays CS_L2
            prm.mode[best] = CS_ATOM;
            prm.rep[best] = rep;
            prm.off[best] = (int)used;
            used += Kb * rep;
        }
    }
    prm.smem_elems = (int)used;
}

template <typename F>
bool index_cat_sparse_fits(int n_cat, const int64_t* K, int64_t p_s) {
    (void)K;
    return n_cat >= 1 && p_s > 0 && p_s < (1ll << 31);
}
template bool index_cat_sparse_fits<float>(int, const int64_t*, int64_t);
template bool index_cat_sparse_fits<double>(int, const int64_t*, int64_t);

template <typename F, int NC>
static int launch_cat_sparse_cols(const F* data, const int32_t* row_idx, const int32_t* indptr,
                                  int p_s, const CodeSrc& rec, const ColOwnerParams& prm, int grid,
                                  cudaStream_t st) {
    const size_t smem = sizeof(F) * ((size_t)prm.slot + 1) * (size_t)prm.cols_per_cta;
    if (smem > 48 * 1024) {
        if (rec.pk)
            TM_CUDA(cudaFuncSetAttribute(k_cat_sparse_cols<F, NC, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            TM_CUDA(cudaFuncSetAttribute(k_cat_sparse_cols<F, NC, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (rec.pk)
        k_cat_sparse_cols<F, NC, true><<<grid, CO_THREADS, smem > 0 ? smem : 16, st>>>(
            data, row_idx, indptr, p_s, rec, prm);
    else
        k_cat_sparse_cols<F, NC, false><<<grid, CO_THREADS, smem > 0 ? smem : 16, st>>>(
            data, row_idx, indptr, p_s, rec, prm);
    TM_LAUNCHED();
    return 0;
}

template <typename F>
static int cat_sparse_cols(const CodeSrc& rec, int n_cat, const int64_t* K, const int32_t* runs,
                           const F* csc_data, const int32_t* csc_row, const int32_t* csc_indptr,
                           int64_t p_s, int n_row_blocks, F* const* outs, F* diag, int64_t diag_ld,
                           cudaStream_t st) {
    ColOwnerParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.n_row_blocks = n_row_blocks;
    prm.diag = diag;
    prm.diag_ld = diag_ld;
    // one resident wave: 4 CTAs of 256 threads per SM
    int grid = sm_count() * 4;
    if (grid > p_s) grid = (int)p_s;
    prm.cols_per_cta = (int)((p_s + grid - 1) / grid);
    grid = (int)((p_s + prm.cols_per_cta - 1) / prm.cols_per_cta);
    // shared-memory column tables: blocks with <= max_k levels while they fit `smem_kb` per CTA
    // (TABMAT_B200_COLS_SMEM_MAXK / TABMAT_B200_COLS_SMEM_KB)
    static const int max_k = getenv("TABMAT_B200_COLS_SMEM_MAXK") ? atoi(getenv("TABMAT_B200_COLS_SMEM_MAXK")) : 512;
    static const int smem_kb = getenv("TABMAT_B200_COLS_SMEM_KB") ? atoi(getenv("TABMAT_B200_COLS_SMEM_KB")) : 40;
    const int64_t budget = (int64_t)(smem_kb * 1024 / sizeof(F)) / prm.cols_per_cta;  // per column
    int64_t slot = 0;
    bool done[IDX_MAX_CATS] = {false};
    for (int it = 0; it < n_cat; ++it) {  // smallest blocks first
        int best = -1;
        for (int c = 0; c < n_cat; ++c)
            if (!done[c] && (best < 0 || K[c] < K[best])) best = c;
        done[best] = true;
        prm.K[best] = (int)K[best];
        prm.runs[best] = runs ? runs[best] : 0;
        prm.rep[best] = 1;
        if (K[best] <= max_k && slot + K[best] <= budget) {
            // replicas (lane % rep) against the CAS-loop contention of shared-memory float
            // atomics (TABMAT_B200_COLS_REP_MAX / _ELEMS: most replicas, largest replicated table)
            static const int rep_max = getenv("TABMAT_B200_COLS_REP_MAX") ? atoi(getenv("TABMAT_B200_COLS_REP_MAX")) : 32;
            static const int rep_elems = getenv("TABMAT_B200_COLS_REP_ELEMS") ? atoi(getenv("TABMAT_B200_COLS_REP_ELEMS")) : 2048;
            int rep = 1;
            while (rep < rep_max && rep < 32 && K[best] * rep * 2 <= rep_elems &&
                   slot + K[best] * rep * 2 <= budget)
                rep *= 2;
            prm.in_smem[best] = 1;
            prm.rep[best] = rep;
            prm.off[best] = (int)slot;
            slot += K[best] * rep;
        }
    }
    prm.slot = (int)slot;
    for (int c = 0; c < n_cat; ++c) {
        prm.out[c] = outs[c];
        if (!prm.in_smem[c])
            TM_CUDA(cudaMemsetAsync(outs[c], 0, sizeof(F) * (size_t)(K[c] * p_s), st));
    }
    switch (n_cat) {
        case 1: return launch_cat_sparse_cols<F, 1>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, grid, st);
        case 2: return launch_cat_sparse_cols<F, 2>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, grid, st);
        case 3: return launch_cat_sparse_cols<F, 3>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, grid, st);
        case 4: return launch_cat_sparse_cols<F, 4>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, grid, st);
        case 5: return launch_cat_sparse_cols<F, 5>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, grid, st);
        case 6: return launch_cat_sparse_cols<F, 6>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, grid, st);
        default:
            return launch_cat_sparse_cols<F, rec_max_cats<F>()>(csc_data, csc_row, csc_indptr, (int)p_s,
                                                                rec, prm, grid, st);
    }
}

// outs[c]: K_c x p_s row-major, overwritten.  n_row_blocks > 1: the CSC arrays are row-blocked
// (blocks of TM_CSC_ROW_BLOCK rows, block slowest, then column, then row).
int index_pack_width(int64_t ncols) {
    int w = 1;
    while ((1ll << w) <= ncols) ++w;   // values 0 .. ncols-1, marker 2^w - 1 = missing
    return w;
}
bool index_pack_fits(int n_cat, const int64_t* K) {
    int bits = 0;
    for (int c = 0; c < n_cat; ++c) bits += index_pack_width(K[c]);
    return n_cat >= 1 && n_cat <= IDX_MAX_CATS && bits <= 64;
}

// rec_v: the 32-byte row records, or NULL when `packed` (codes of every CSC non-zero's row,
// bit-packed block after block from bit 0, widths index_pack_width(K_c)) and `d` are given.
template <typename F>
int index_cat_sparse(const void* rec_v, const F* d, const uint64_t* packed, int n_cat,
                     const int64_t* K, const int32_t* runs, const F* csc_data,
                     const int32_t* csc_row, const int32_t* csc_indptr, int64_t p_s,
                     int n_row_blocks, F* const* outs, F* diag, int64_t diag_ld, cudaStream_t st) {
    CodeSrc rec;
    memset(&rec, 0, sizeof(rec));
    rec.rec = rec_v;
    if (!rec_v) {
        if (!packed || !d || !index_pack_fits(n_cat, K))
            return fail("index_cat_sparse: neither row records nor packed codes");
        rec.d = d;
        rec.pk = reinterpret_cast<const unsigned long long*>(packed);
        int sh = 0;
        for (int c = 0; c < n_cat; ++c) {
            const int w = index_pack_width(K[c]);
            rec.shift[c] = sh;
            rec.mask[c] = w >= 32 ? 0xffffffffu : ((1u << w) - 1u);
            sh += w;
        }
    }
    if (n_row_blocks > 1)
        return cat_sparse_cols<F>(rec, n_cat, K, runs, csc_data, csc_row, csc_indptr, p_s,
                                  n_row_blocks, outs, diag, diag_ld, st);
    if (diag) return fail("index_cat_sparse: the diagonal by-product needs the row-blocked CSC");
    CatSparseParams prm;
    cat_sparse_layout<F>(n_cat, K, runs, prm);
    for (int c = 0; c < n_cat; ++c) {
        prm.out[c] = outs[c];
        if (prm.mode[c] == CS_L2)
            TM_CUDA(cudaMemsetAsync(outs[c], 0, sizeof(F) * (size_t)(K[c] * p_s), st));
    }
    switch (n_cat) {
        case 1: return launch_cat_sparse<F, 1>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, st);
        case 2: return launch_cat_sparse<F, 2>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, st);
        case 3: return launch_cat_sparse<F, 3>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, st);
        case 4: return launch_cat_sparse<F, 4>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, st);
        case 5: return launch_cat_sparse<F, 5>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, st);
        case 6: return launch_cat_sparse<F, 6>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, st);
        default:
            return launch_cat_sparse<F, rec_max_cats<F>()>(csc_data, csc_row, csc_indptr, (int)p_s,
                                                           rec, prm, st);
    }
}
template int index_cat_sparse<float>(const void*, const float*, const uint64_t*, int,
                                     const int64_t*, const int32_t*, const float*, const int32_t*,
                                     const int32_t*, int64_t, int, float* const*, float*, int64_t,
                                     cudaStream_t);
template int index_cat_sparse<double>(const void*, const double*, const uint64_t*, int,
                                      const int64_t*, const int32_t*, const double*,
                                      const int32_t*, const int32_t*, int64_t, int,
                                      double* const*, double*, int64_t, cudaStream_t);

}  // namespace tmb
