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
//                      of the sparse block: a CTA owns G columns j, walks their non-zeros,
//                      gathers the 32-byte record of each row and accumulates
//                      d[k] * A[k, j] into shared-memory columns out_i[:, j]; the columns are
//                      then stored once — no global atomics at all, HBM-bound
//                      (8 bytes per non-zero streamed + one sector per non-zero gathered).
#include <cstdlib>

#include "tm_common.cuh"

namespace tmb {

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

template <typename F>
__device__ __forceinline__ void pair_add(const PairParams& prm, F* smem, int t, long long key,
                                         long long size, F val, bool runs, int lane) {
    bool head = true;
    if (runs) head = run_reduce<F, long long>(key, val, lane);
    if (!head || key < 0) return;
    if (prm.smem_off[t] >= 0)
        atomicAdd(smem + prm.smem_off[t] + (long long)(lane % prm.copies[t]) * size + key, val);
    else
        red_add(static_cast<F*>(prm.out[t]) + key, val);
}

template <typename F, int NC, int U>
__global__ void __launch_bounds__(1024, 1)
k_cat_pairs(const RowRec<F>* __restrict__ rec, int64_t n, const PairParams prm) {
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
            if (k < n) {
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
struct CatSparseParams {
    int K[IDX_MAX_CATS];
    int rep[IDX_MAX_CATS];       // replicas of block c's column table (lane % rep)
    int off[IDX_MAX_CATS + 1];   // element offset of block c's table inside one column slot
    int runs[IDX_MAX_CATS];
    void* out[IDX_MAX_CATS];     // K_c x p_s, row-major; every element is written
    int G;                       // sparse columns per CTA visit
};

template <typename F, int NC, int U>
__global__ void __launch_bounds__(512)
k_cat_sparse_csc(const F* __restrict__ data, const int32_t* __restrict__ row_idx,
                 const int32_t* __restrict__ indptr, int p_s,
                 const RowRec<F>* __restrict__ rec, const CatSparseParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    F* smem = reinterpret_cast<F*>(smem_raw);
    const int slot = prm.off[NC];  // elements per column slot
    const int G = prm.G;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int nwib = blockDim.x >> 5;
    const int n_groups = (p_s + G - 1) / G;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int j0 = grp * G;
        const int gc = min(G, p_s - j0);
        for (int i = threadIdx.x; i < gc * slot; i += blockDim.x) smem[i] = F(0);
        __syncthreads();
        for (int g = 0; g < gc; ++g) {
            F* tab = smem + g * slot;
            const int e0 = indptr[j0 + g], e1 = indptr[j0 + g + 1];
            for (int eb = e0 + wib * (32 * U); eb < e1; eb += nwib * (32 * U)) {
                RowRec<F> r[U];
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
                        r[u] = load_rec<F>(rec, k);
                    } else {
                        r[u].d = F(0);
#pragma unroll
                        for (int c = 0; c < rec_max_cats<F>(); ++c) r[u].c[c] = -1;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const F val0 = r[u].d * a[u];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        int key = r[u].c[c];
                        F val = val0;
                        bool head = true;
                        if (prm.runs[c]) head = run_reduce<F, int>(key, val, lane);
                        if (head && key >= 0)
                            atomicAdd(tab + prm.off[c] + (lane % prm.rep[c]) * prm.K[c] + key, val);
                    }
                }
            }
        }
        __syncthreads();
        // store the gc columns: g fastest so that a thread group writes consecutive j
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            F* out = static_cast<F*>(prm.out[c]);
            const int Kc = prm.K[c], rep = prm.rep[c];
            for (int i = threadIdx.x; i < Kc * gc; i += blockDim.x) {
                const int lvl = i / gc, g = i - lvl * gc;
                const F* tab = smem + g * slot + prm.off[c] + lvl;
                F s = F(0);
                for (int q = 0; q < rep; ++q) s += tab[q * Kc];
                out[(int64_t)lvl * p_s + j0 + g] = s;
            }
        }
        __syncthreads();
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

constexpr size_t IDX_SMEM_BUDGET = 160 * 1024;

template <typename F, int NC>
static int launch_pairs(const RowRec<F>* rec, int64_t n, const PairParams& prm, cudaStream_t st) {
    constexpr int U = 2;
    const size_t smem = sizeof(F) * (size_t)prm.smem_elems;
    static bool attr_set = false;
    if (!attr_set) {
        TM_CUDA(cudaFuncSetAttribute(k_cat_pairs<F, NC, U>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)IDX_SMEM_BUDGET));
        attr_set = true;
    }
    const int g = grid_for(n, 1024 * U, sm_count());
    k_cat_pairs<F, NC, U><<<g, 1024, smem, st>>>(rec, n, prm);
    TM_LAUNCHED();
    return 0;
}

// outs_self[c]: K_c values; outs_pair[i * n_cat + j] (i < j): K_i x K_j row-major.  Overwrites.
template <typename F>
int index_cat_pairs(const void* rec_v, int64_t n, int n_cat, const int64_t* K, const int32_t* runs,
                    F* const* outs_self, F* const* outs_pair, cudaStream_t st) {
    const RowRec<F>* rec = static_cast<const RowRec<F>*>(rec_v);
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
    // shared-memory placement: smallest tables first, tiny ones replicated (up to one per lane)
    size_t budget = IDX_SMEM_BUDGET / sizeof(F);
    size_t used = 0;
    bool placed[IDX_MAX_TARGETS] = {false};
    for (;;) {
        int best = -1;
        for (int t = 0; t < nt; ++t)
            if (!placed[t] && (best < 0 || size[t] < size[best])) best = t;
        if (best < 0) break;
        placed[best] = true;
        if ((size_t)size[best] > budget - used) continue;  // larger ones will not fit either
        int copies = 1;
        while (copies < 32 && (int64_t)size[best] * copies * 2 <= 1024) copies *= 2;
        if ((size_t)size[best] * copies > budget - used) copies = 1;
        prm.smem_off[best] = (int)used;
        prm.copies[best] = copies;
        used += (size_t)size[best] * copies;
    }
    prm.smem_elems = (int)used;
    switch (n_cat) {
        case 1: return launch_pairs<F, 1>(rec, n, prm, st);
        case 2: return launch_pairs<F, 2>(rec, n, prm, st);
        case 3: return launch_pairs<F, 3>(rec, n, prm, st);
        case 4: return launch_pairs<F, 4>(rec, n, prm, st);
        case 5: return launch_pairs<F, 5>(rec, n, prm, st);
        case 6: return launch_pairs<F, 6>(rec, n, prm, st);
        default: return launch_pairs<F, rec_max_cats<F>()>(rec, n, prm, st);
    }
}
template int index_cat_pairs<float>(const void*, int64_t, int, const int64_t*, const int32_t*,
                                    float* const*, float* const*, cudaStream_t);
template int index_cat_pairs<double>(const void*, int64_t, int, const int64_t*, const int32_t*,
                                     double* const*, double* const*, cudaStream_t);

template <typename F, int NC>
static int launch_cat_sparse(const F* data, const int32_t* row_idx, const int32_t* indptr,
                             int p_s, const RowRec<F>* rec, const CatSparseParams& prm,
                             size_t smem, cudaStream_t st) {
    constexpr int U = 2;
    static bool attr_set = false;
    if (!attr_set) {
        TM_CUDA(cudaFuncSetAttribute(k_cat_sparse_csc<F, NC, U>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)IDX_SMEM_BUDGET));
        attr_set = true;
    }
    const int n_groups = (p_s + prm.G - 1) / prm.G;
    const int g = n_groups < sm_count() * 4 ? n_groups : sm_count() * 4;
    k_cat_sparse_csc<F, NC, U><<<g, 512, smem, st>>>(data, row_idx, indptr, p_s, rec, prm);
    TM_LAUNCHED();
    return 0;
}

// can the column tables of all blocks (with their replicas) live in shared memory?
template <typename F>
static bool cat_sparse_layout(int n_cat, const int64_t* K, const int32_t* runs, int64_t p_s,
                              CatSparseParams& prm) {
    memset(&prm, 0, sizeof(prm));
    int64_t slot = 0;
    for (int c = 0; c < n_cat; ++c) {
        int rep = 1;
        while (rep < 8 && K[c] * rep * 2 <= 2048) rep *= 2;
        prm.K[c] = (int)K[c];
        prm.rep[c] = rep;
        prm.runs[c] = runs ? runs[c] : 0;
        prm.off[c] = (int)slot;
        slot += K[c] * rep;
    }
    prm.off[n_cat] = (int)slot;
    const int64_t budget = (int64_t)(IDX_SMEM_BUDGET / sizeof(F));
    if (slot > budget) return false;
    int64_t G = budget / slot;
    if (G > 8) G = 8;
    // enough CTAs to fill the machine a few times over
    while (G > 1 && (p_s + G - 1) / G < (int64_t)sm_count() * 4) --G;
    prm.G = (int)G;
    return true;
}

template <typename F>
bool index_cat_sparse_fits(int n_cat, const int64_t* K, int64_t p_s) {
    CatSparseParams prm;
    return p_s > 0 && p_s < (1ll << 31) && cat_sparse_layout<F>(n_cat, K, nullptr, p_s, prm);
}
template bool index_cat_sparse_fits<float>(int, const int64_t*, int64_t);
template bool index_cat_sparse_fits<double>(int, const int64_t*, int64_t);

// outs[c]: K_c x p_s row-major, fully overwritten.
template <typename F>
int index_cat_sparse(const void* rec_v, int n_cat, const int64_t* K, const int32_t* runs,
                     const F* csc_data, const int32_t* csc_row, const int32_t* csc_indptr,
                     int64_t p_s, F* const* outs, cudaStream_t st) {
    const RowRec<F>* rec = static_cast<const RowRec<F>*>(rec_v);
    CatSparseParams prm;
    if (!cat_sparse_layout<F>(n_cat, K, runs, p_s, prm))
        return fail("index_cat_sparse: column tables exceed shared memory");
    for (int c = 0; c < n_cat; ++c) prm.out[c] = outs[c];
    const size_t smem = sizeof(F) * (size_t)prm.off[n_cat] * (size_t)prm.G;
    switch (n_cat) {
        case 1: return launch_cat_sparse<F, 1>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, smem, st);
        case 2: return launch_cat_sparse<F, 2>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, smem, st);
        case 3: return launch_cat_sparse<F, 3>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, smem, st);
        case 4: return launch_cat_sparse<F, 4>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, smem, st);
        case 5: return launch_cat_sparse<F, 5>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, smem, st);
        case 6: return launch_cat_sparse<F, 6>(csc_data, csc_row, csc_indptr, (int)p_s, rec, prm, smem, st);
        default:
            return launch_cat_sparse<F, rec_max_cats<F>()>(csc_data, csc_row, csc_indptr, (int)p_s, rec,
                                                           prm, smem, st);
    }
}
template int index_cat_sparse<float>(const void*, int, const int64_t*, const int32_t*, const float*,
                                     const int32_t*, const int32_t*, int64_t, float* const*,
                                     cudaStream_t);
template int index_cat_sparse<double>(const void*, int, const int64_t*, const int32_t*,
                                      const double*, const int32_t*, const int32_t*, int64_t,
                                      double* const*, cudaStream_t);

}  // namespace tmb
