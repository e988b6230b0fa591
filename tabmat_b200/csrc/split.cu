// SplitMatrix.sandwich as ONE native call (reference: the Python block loop of
// split_matrix.py:324-356, which makes one native call per self block and per cross pair and
// assembles the result with numpy fancy indexing).
//
//   tm_split_sandwich_blocks_*   computes every self block and every cross block of the column
//                                blocks into one flat workspace (layout below);
//   tm_split_sandwich_assemble_* places the workspace into the p x p float64 result
//                                (split_matrix.py:336-354) — kept separate so that a row-sharded
//                                multi-GPU caller can allreduce the flat workspace in between.
//
// Workspace layout (elements of the block dtype), blocks in the given order:
//   for i in 0..nb-1:  self_i   dense / sparse: ncols_i x ncols_i (row-major, symmetric)
//                               categorical:    ncols_i           (the diagonal)
//     for j in i+1..nb-1: cross_ij, stored as (rows of block a) x (cols of block b), row-major,
//                         where (a, b) is the kernel's native orientation:
//                           categorical x dense / sparse x dense / categorical x sparse /
//                           categorical_i x categorical_j (i < j)
#include <climits>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "tm_common.cuh"

namespace tmb {

enum { KIND_DENSE = 0, KIND_SPARSE = 1, KIND_CAT = 2 };

static inline int64_t self_elems(const tm_block_desc& b) {
    return b.kind == KIND_CAT ? b.ncols : b.ncols * b.ncols;
}
// every block of the workspace starts at a multiple of 4 elements: the scatter kernels write
// 16-byte vector REDs (an odd-width categorical block in front of the dense block used to leave
// the dense cross blocks misaligned)
static inline int64_t ws_pad(int64_t elems) { return (elems + 3) & ~int64_t(3); }

// orientation of the stored cross block of blocks (i, j), i < j: returns true when the stored
// rows belong to block j (i.e. the kernel's native orientation is (j, i))
static inline bool cross_rows_are_j(const tm_block_desc& bi, const tm_block_desc& bj) {
    if (bi.kind == KIND_DENSE) return true;                           // (sparse|cat) x dense
    if (bi.kind == KIND_SPARSE && bj.kind == KIND_CAT) return true;   // cat x sparse
    return false;  // sparse x dense (i sparse, j dense), cat x dense, cat x sparse, cat_i x cat_j
}

// overload shims over the extern "C" entry points
#define TM_SHIM(name, F, SUF)                                                     \
    template <typename... A>                                                      \
    static inline int name(F*, A... a) {                                          \
        return tm_##name##_##SUF(a...);                                           \
    }
TM_SHIM(dense_sandwich, float, f32)
TM_SHIM(dense_sandwich, double, f64)
TM_SHIM(sparse_sandwich, float, f32)
TM_SHIM(sparse_sandwich, double, f64)
TM_SHIM(csr_dense_sandwich, float, f32)
TM_SHIM(csr_dense_sandwich, double, f64)
TM_SHIM(cat_sandwich, float, f32)
TM_SHIM(cat_sandwich, double, f64)
TM_SHIM(cat_dense_sandwich, float, f32)
TM_SHIM(cat_dense_sandwich, double, f64)
TM_SHIM(cat_cat_sandwich, float, f32)
TM_SHIM(cat_cat_sandwich, double, f64)
TM_SHIM(cat_sparse_sandwich, float, f32)
TM_SHIM(cat_sparse_sandwich, double, f64)
TM_SHIM(scatter_block, float, f32)
TM_SHIM(scatter_block, double, f64)
TM_SHIM(scatter_diag, float, f32)
TM_SHIM(scatter_diag, double, f64)

// two non-blocking side streams + four events per thread (created lazily, never destroyed);
// TABMAT_B200_SIDE_STREAM=0 keeps everything on the caller's stream
static cudaStream_t side_stream(int which) {
    static thread_local cudaStream_t st[2] = {nullptr, nullptr};
    static thread_local bool tried = false;
    if (!tried) {
        tried = true;
        const char* e = getenv("TABMAT_B200_SIDE_STREAM");
        if (!(e && atoi(e) == 0)) {
            cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking);
            cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking);
        }
    }
    return st[which];
}
static cudaEvent_t side_event(int i) {
    static thread_local cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (!ev[i]) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    return ev[i];
}
// pass schedule of tm_split_sandwich_blocks_* (see the switch at the end of split_blocks)
static int initial_sched() {
    const char* e = getenv("TABMAT_B200_SCHED");
    if (e) return atoi(e);
    const char* sf = getenv("TABMAT_B200_SCATTER_FIRST");
    // default: serial.  Measured on B200 (n = 4e7): every schedule lands within 1 ms of the
    // sum of the three passes timed alone (28.0-29.9 ms) — they all queue on the same L2
    // slices — so the overlap buys nothing and the serial order keeps the pass timings clean.
    return (sf && atoi(sf) == 1) ? 1 : 3;
}
static int g_split_sched = initial_sched();
// which forms the last tm_split_sandwich_blocks_* call used (tm_split_last_plan): bit 0 = the
// tcgen05 pass ran, bit 1 = categorical scatter work inside it, bit 2 = sparse scatter work inside
// it, bit 3 = dense x sparse by gather
static int g_last_plan = 0;

// ---- optional pass-level timing (tm_split_profile_*): CUDA events on the stream each pass
// is launched on; bench.py reads them after synchronising ---------------------------------
enum { PASS_TENSOR = 0, PASS_SCATTER = 1, PASS_INDEX = 2, PASS_COUNT = 3 };
constexpr int PROF_RING = 64;  // calls kept since tm_split_profile_enable(1)
static bool g_profile = false;
static int g_prof_calls = 0;
static cudaEvent_t g_pass_ev[PROF_RING][PASS_COUNT][2];
static bool g_pass_used[PROF_RING][PASS_COUNT];
// NVTX range per pass (visible in Nsight Systems / ncu --nvtx; a no-op without a tool attached)
static const char* const kPassName[PASS_COUNT] = {"tabmat_b200:tensor_pass",
                                                  "tabmat_b200:scatter_pass",
                                                  "tabmat_b200:index_pass"};
static void pass_mark(int pass, int which, cudaStream_t st) {
    if (which == 0)
        nvtxRangePushA(kPassName[pass]);
    else
        nvtxRangePop();
    if (!g_profile) return;
    const int slot = g_prof_calls % PROF_RING;
    if (!g_pass_ev[slot][pass][which]) cudaEventCreate(&g_pass_ev[slot][pass][which]);
    cudaEventRecord(g_pass_ev[slot][pass][which], st);
    if (which == 1) g_pass_used[slot][pass] = true;
}

static inline int64_t n_rows_or_all(const int32_t* rows, int64_t n_rows, int64_t n) {
    return rows ? n_rows : n;
}

TM_SHIM(dense_rmatvec, float, f32)
TM_SHIM(dense_rmatvec, double, f64)

// `v` / `dense_vec` (fused IRLS pass, both or neither): dense_vec[c] = sum_k v[k] X_dense[k, c]
// over `rows`, computed by the same pass over the dense block as its sandwich (the tcgen05
// kernel's scale warps, fp32 FMAs) or, where that kernel does not apply, by the GEMV kernel.
template <typename F>
int split_blocks(const tm_block_desc* blk, int nb, int64_t n, const F* d, const int32_t* rows,
                 int64_t n_rows, F* ws, tm_stream_t stream, int part = 0, const F* v = nullptr,
                 F* dense_vec = nullptr) {
    F* tag = nullptr;
    if (nb <= 0) return 0;
    if (nb > 64) return fail("tm_split_sandwich: more than 64 blocks");
    // offsets
    std::vector<int64_t> self_off(nb);
    std::vector<std::vector<int64_t>> cross_off(nb, std::vector<int64_t>(nb, -1));
    int64_t off = 0;
    for (int i = 0; i < nb; ++i) {
        self_off[i] = off;
        off += ws_pad(self_elems(blk[i]));
        for (int j = i + 1; j < nb; ++j) {
            cross_off[i][j] = off;
            off += ws_pad(blk[i].ncols * blk[j].ncols);
        }
    }
    // can the dense-operand cross blocks be fused into one pass?
    int dense_idx = -1, n_dense = 0, n_sparse = 0, n_cat = 0, sparse_idx = -1;
    for (int i = 0; i < nb; ++i) {
        if (blk[i].kind == KIND_DENSE) { dense_idx = i; ++n_dense; }
        else if (blk[i].kind == KIND_SPARSE) { sparse_idx = i; ++n_sparse; }
        else if (blk[i].kind == KIND_CAT) ++n_cat;
        else return fail("tm_split_sandwich: unknown block kind");
    }
    constexpr int W = sizeof(F) == 4 ? 4 : 2;
    bool fuse = n_dense == 1 && n_sparse <= 1 && n_cat <= 8 && (n_cat + n_sparse) > 0;
    if (fuse) {
        const tm_block_desc& D = blk[dense_idx];
        fuse = D.c_order && D.ncols % W == 0 && D.ncols <= 64 * W &&
               (reinterpret_cast<uintptr_t>(D.data) & 15) == 0;
    }
    // ---- plan of the tensor pass (host-side decisions only; launched by tensor_pass below) ----
    // tcgen05 weighted SYRK of the dense block + one-hot MMAs for the few-level categoricals,
    // and the opt-in sorted-gather kernel for categorical blocks that carry a row permutation.
    std::vector<char> on_tensor(nb, 0);
    bool dense_self_done = false;
    bool tc_ok = false, any_gather = false;
    TcOneHot oh;
    oh.ncat = 0;
    int64_t oh_slots = 0;
    int oh_which[8];
    const int64_t nr_all = n_rows_or_all(rows, n_rows, n);
    if (fuse && sizeof(F) == 4) {
        const tm_block_desc& D = blk[dense_idx];
        static const bool onehot_off =
            getenv("TABMAT_B200_ONEHOT") && atoi(getenv("TABMAT_B200_ONEHOT")) == 0;
        static const bool gather_off =
            getenv("TABMAT_B200_GATHER") && atoi(getenv("TABMAT_B200_GATHER")) == 0;
        tc_ok = g_dense_f32_mode != 1 && nr_all > 0 && dense_tc_eligible(n, D.ncols, 1, D.data);
        if (tc_ok) {
            for (int i = 0; i < nb && oh.ncat < 8 && D.ncols <= 128 && !onehot_off; ++i) {
                if (blk[i].kind != KIND_CAT || blk[i].ncols <= 0 || blk[i].ncols > 256) continue;
                if (oh_slots + blk[i].ncols > TC_ONEHOT_MAX_SLOTS) continue;
                oh_which[oh.ncat] = i;
                oh.codes[oh.ncat] = static_cast<const int32_t*>(blk[i].data);
                oh.K[oh.ncat] = (int)blk[i].ncols;
                oh.drop_first[oh.ncat] = blk[i].drop_first;
                oh_slots += blk[i].ncols;
                on_tensor[i] = 1;
                ++oh.ncat;
            }
            dense_self_done = true;
        }
        for (int i = 0; i < nb && !gather_off; ++i)
            if (blk[i].kind == KIND_CAT && !on_tensor[i] && blk[i].cat_perm) {
                on_tensor[i] = 2;
                any_gather = true;
            }
    }
    const bool have_tensor = tc_ok || any_gather;
    // dense x sparse by row-blocked gather instead of one RED per non-zero (TABMAT_B200_DXS:
    // "gather" (default when the matrix carries the blocked copy) | "red")
    static const bool gather_env_off = getenv("TABMAT_B200_DXS") && !strcmp(getenv("TABMAT_B200_DXS"), "red");
    const bool sparse_by_gather = fuse && !gather_env_off && sparse_idx >= 0 &&
                                  blk[sparse_idx].nnz > 0 && blk[sparse_idx].gcsc_data &&
                                  blk[sparse_idx].gcsc_indices && blk[sparse_idx].gcsc_indptr &&
                                  blk[sparse_idx].gcsc_row_blocks > 0;

    // fused form: the scatter work (dense x many-level categoricals, dense x sparse) rides along
    // the tcgen05 kernel as extra warps reading the TMA-staged tile, so X is read once
    int n_scatter_cats = 0;
    for (int i = 0; i < nb; ++i)
        if (blk[i].kind == KIND_CAT && !on_tensor[i]) ++n_scatter_cats;
    const bool sparse_for_tc = sparse_idx >= 0 && blk[sparse_idx].nnz > 0 && !sparse_by_gather;
    const bool scatter_in_tc =
        tc_ok && fuse &&
        dense_tc_scatter_eligible(blk[dense_idx].ncols, n_scatter_cats, sparse_for_tc) &&
        (n_scatter_cats > 0 || sparse_for_tc);

    g_last_plan = (tc_ok ? 1 : 0) | (scatter_in_tc && n_scatter_cats > 0 ? 2 : 0) |
                  (scatter_in_tc && sparse_for_tc ? 4 : 0) | (sparse_by_gather ? 8 : 0);
    // destinations and sources of the scatter work (shared by the fused tensor pass and the
    // stand-alone scatter pass)
    struct ScatterPlan {
        const int32_t* codes[8];
        int64_t K[8];
        int32_t df[8];
        F* outs[8];
        int c = 0;
        const F* sdata = nullptr;
        const int32_t *sind = nullptr, *sptr = nullptr;
        int64_t ps = 0;
        F* out_s = nullptr;
    } sp;
    if (fuse) {
        for (int i = 0; i < nb; ++i) {
            if (blk[i].kind != KIND_CAT || on_tensor[i]) continue;
            sp.codes[sp.c] = static_cast<const int32_t*>(blk[i].data);
            sp.K[sp.c] = blk[i].ncols;
            sp.df[sp.c] = blk[i].drop_first;
            int a = i < dense_idx ? i : dense_idx, b = i < dense_idx ? dense_idx : i;
            sp.outs[sp.c] = ws + cross_off[a][b];
            ++sp.c;
        }
        if (sparse_idx >= 0 && blk[sparse_idx].nnz > 0) {
            const tm_block_desc& S = blk[sparse_idx];
            sp.sdata = static_cast<const F*>(S.data);
            sp.sind = S.csr_indices;
            sp.sptr = S.csr_indptr;
            sp.ps = S.ncols;
            int a = sparse_idx < dense_idx ? sparse_idx : dense_idx;
            int b = sparse_idx < dense_idx ? dense_idx : sparse_idx;
            sp.out_s = ws + cross_off[a][b];
        }
    }

    auto tensor_pass = [&](cudaStream_t st, bool share_sm) -> int {
        if (!have_tensor) return 0;
        const tm_block_desc& D = blk[dense_idx];
        pass_mark(PASS_TENSOR, 0, st);
        Scratch dm(rows ? sizeof(float) * (size_t)n : 0, st);
        Scratch vm(rows && v && tc_ok ? sizeof(float) * (size_t)n : 0, st);
        if (dm.err != cudaSuccess) return fail_cuda(dm.err, "scratch");
        if (vm.err != cudaSuccess) return fail_cuda(vm.err, "scratch");
        const float* dd = reinterpret_cast<const float*>(d);
        const float* vv = reinterpret_cast<const float*>(v);
        if (rows) {
            int rc = masked_weights<float>(dd, n, rows, n_rows, dm.as<float>(), st);
            if (rc) return rc;
            dd = dm.as<float>();
            if (vv && tc_ok) {
                rc = masked_weights<float>(vv, n, rows, n_rows, vm.as<float>(), st);
                if (rc) return rc;
                vv = vm.as<float>();
            }
        }
        if (tc_ok) {
            Scratch tmp(sizeof(float) * (size_t)(oh_slots > 0 ? oh_slots : 1) * (size_t)D.ncols, st);
            if (tmp.err != cudaSuccess) return fail_cuda(tmp.err, "scratch");
            oh.out = tmp.as<float>();
            FusedCrossParams fc;
            CrossScratch fscr;
            if (scatter_in_tc) {
                const bool sp_here = !sparse_by_gather;   // else the gather kernel owns that block
                int rc = cross_prepare<float>(
                    D.ncols, sp.c, sp.codes, sp.K, sp.df, reinterpret_cast<float* const*>(sp.outs),
                    sp_here ? reinterpret_cast<const float*>(sp.sdata) : nullptr,
                    sp_here ? sp.sind : nullptr, sp_here ? sp.sptr : nullptr, sp_here ? sp.ps : 0,
                    sp_here ? reinterpret_cast<float*>(sp.out_s) : nullptr, fc, fscr, st);
                if (rc) return rc;
            }
            int rc = dense_sandwich_tc_f32(static_cast<const float*>(D.data), n, D.ncols, 1, dd,
                                           reinterpret_cast<float*>(ws + self_off[dense_idx]), st,
                                           oh.ncat ? &oh : nullptr, share_sm,
                                           scatter_in_tc ? &fc : nullptr, vv,
                                           reinterpret_cast<float*>(dense_vec));
            if (rc) return rc;
            if (scatter_in_tc) {
                rc = cross_finish<float>(D.ncols, sp.c, sp.K,
                                         reinterpret_cast<float* const*>(sp.outs), fc, st);
                if (rc) return rc;
            }
            int64_t o = 0;
            for (int c = 0; c < oh.ncat; ++c) {
                int i = oh_which[c];
                int a = i < dense_idx ? i : dense_idx, b = i < dense_idx ? dense_idx : i;
                TM_CUDA(cudaMemcpyAsync(ws + cross_off[a][b], tmp.as<float>() + o * D.ncols,
                                        sizeof(float) * (size_t)(blk[i].ncols * D.ncols),
                                        cudaMemcpyDeviceToDevice, st));
                o += blk[i].ncols;
            }
        }
        // categorical blocks with a sorted row permutation: HBM-bound gather
        for (int i = 0; i < nb; ++i) {
            if (on_tensor[i] != 2) continue;
            int a = i < dense_idx ? i : dense_idx, b = i < dense_idx ? dense_idx : i;
            int rc = cat_dense_gather_f32(static_cast<const float*>(D.data), D.ncols, dd,
                                          blk[i].cat_perm, blk[i].cat_segptr, blk[i].ncols,
                                          blk[i].cat_nvalid,
                                          reinterpret_cast<float*>(ws + cross_off[a][b]), st);
            if (rc) return rc;
        }
        pass_mark(PASS_TENSOR, 1, st);
        return 0;
    };

    // ---- scatter pass: dense x many-level categoricals + dense x sparse (vector REDs) -------
    auto scatter_pass = [&](cudaStream_t st) -> int {
        if (!fuse) return 0;
        const tm_block_desc& D = blk[dense_idx];
        if (sparse_idx >= 0 && !sp.out_s) {  // empty sparse block: its cross block is zero
            int a = sparse_idx < dense_idx ? sparse_idx : dense_idx;
            int b = sparse_idx < dense_idx ? dense_idx : sparse_idx;
            TM_CUDA(cudaMemsetAsync(ws + cross_off[a][b], 0,
                                    sizeof(F) * (size_t)(blk[a].ncols * blk[b].ncols), st));
        }
        if (scatter_in_tc && !sparse_by_gather) return 0;  // done by the tcgen05 kernel's scatter warps
        if (sp.c > 0 || sp.out_s) {
            pass_mark(PASS_SCATTER, 0, st);
            int rc = 0;
            if (sparse_by_gather) {
                // dense x sparse: row-blocked gather (one RED per (row block, column) run)
                const tm_block_desc& S = blk[sparse_idx];
                Scratch dmg(rows ? sizeof(F) * (size_t)n : 0, st);
                if (dmg.err != cudaSuccess) return fail_cuda(dmg.err, "scratch");
                const F* dd = d;
                if (rows) {
                    rc = masked_weights<F>(d, n, rows, n_rows, dmg.as<F>(), st);
                    if (rc) return rc;
                    dd = dmg.as<F>();
                }
                rc = csc_dense_gather<F>(static_cast<const F*>(D.data), D.ncols, dd,
                                         static_cast<const F*>(S.gcsc_data), S.gcsc_indices,
                                         S.gcsc_indptr, S.ncols, S.gcsc_row_blocks, sp.out_s, st);
                if (rc) return rc;
                // dense x many-level categoricals: the scatter warps of the tcgen05 kernel when it
                // runs in its fused form, else the RED kernel without a sparse operand
                if (sp.c > 0 && !scatter_in_tc)
                    rc = dense_cross_fused<F>(static_cast<const F*>(D.data), n, D.ncols, d, rows,
                                              n_rows, sp.c, sp.codes, sp.K, sp.df, sp.outs,
                                              (const F*)nullptr, nullptr, nullptr, 0, (F*)nullptr,
                                              /*runs=*/1, st);
            } else {
                rc = dense_cross_fused<F>(static_cast<const F*>(D.data), n, D.ncols, d, rows,
                                          n_rows, sp.c, sp.codes, sp.K, sp.df, sp.outs, sp.sdata,
                                          sp.sind, sp.sptr, sp.ps, sp.out_s, /*runs=*/1, st);
            }
            if (rc) return rc;
            pass_mark(PASS_SCATTER, 1, st);
        }
        return 0;
    };

    // ---- index pass: every block without the dense operand -----------------------------------
    auto index_pass = [&](cudaStream_t st) -> int {
        tm_stream_t stream = reinterpret_cast<tm_stream_t>(st);
        pass_mark(PASS_INDEX, 0, st);
        // fused index blocks (split_index.cu): all categorical self / pair blocks in one pass
        // over 32-byte row records, categorical x sparse for all categorical blocks from the
        // CSC copy without global atomics
        bool cats_fused = false, cat_sparse_fused = false, sparse_diag_done = false;
        cudaStream_t sparse_side = nullptr;
        {
            int cats[8];
            int nc = 0;
            for (int i = 0; i < nb; ++i)
                if (blk[i].kind == KIND_CAT) {
                    if (nc < 8) cats[nc] = i;
                    ++nc;
                }
            int64_t Kc[8];
            const int32_t* cc[8];
            int32_t dfc[8], runc[8], runp[8];
            bool ok = nc >= 1 && nc <= 7 && nr_all > 0 && n > 0;
            if (ok) {
                for (int a = 0; a < nc; ++a) {
                    const tm_block_desc& b = blk[cats[a]];
                    Kc[a] = b.ncols;
                    cc[a] = static_cast<const int32_t*>(b.data);
                    dfc[a] = b.drop_first;
                    runc[a] = (b.flags & TM_BLOCK_FLAG_RUNS) ? 1 : 0;
                    // down a CSC column the rows are ~n/nnz_col apart: only the primary sort
                    // key still forms runs there
                    runp[a] = (b.flags & TM_BLOCK_FLAG_PRIMARY) ? 1 : 0;
                }
                ok = index_fused_eligible<F>(nc, Kc);
            }
            // categorical x sparse from the CSC copy: with the bit-packed codes of the
            // non-zeros' rows (built once per matrix) nothing is packed per call; without them
            // the 32-byte row records {d, codes} are built first and gathered per non-zero
            const bool want_cs = ok && sparse_idx >= 0 && blk[sparse_idx].csc_indptr &&
                                 blk[sparse_idx].csc_indices && blk[sparse_idx].csc_data &&
                                 index_cat_sparse_fits<F>(nc, Kc, blk[sparse_idx].ncols);
            static const bool packed_off =
                getenv("TABMAT_B200_CSC_PACKED") && atoi(getenv("TABMAT_B200_CSC_PACKED")) == 0;
            const bool use_packed = want_cs && !packed_off && blk[sparse_idx].csc_cat_codes &&
                                    index_pack_fits(nc, Kc);
            // TABMAT_B200_PAIRS_DIRECT=0: the pairs kernel reads the packed row records too
            static const bool pairs_direct_off = getenv("TABMAT_B200_PAIRS_DIRECT") &&
                                                 atoi(getenv("TABMAT_B200_PAIRS_DIRECT")) == 0;
            const bool need_rec = (want_cs && !use_packed) || (ok && pairs_direct_off);
            Scratch rec(need_rec ? index_record_bytes<F>(n) : 0, st);
            Scratch dmi(ok && rows ? sizeof(F) * (size_t)n : 0, st);
            if (rec.err != cudaSuccess) return fail_cuda(rec.err, "scratch");
            if (dmi.err != cudaSuccess) return fail_cuda(dmi.err, "scratch");
            if (ok) {
                const F* dd = d;
                if (rows) {
                    int rc = masked_weights<F>(d, n, rows, n_rows, dmi.as<F>(), st);
                    if (rc) return rc;
                    dd = dmi.as<F>();
                }
                int rc = 0;
                if (need_rec) {
                    rc = index_pack_records<F>(dd, n, nc, cc, dfc, rec.p, st);
                    if (rc) return rc;
                }
                F* outs_self[8];
                F* outs_pair[64];
                for (int a = 0; a < nc; ++a) {
                    outs_self[a] = ws + self_off[cats[a]];
                    for (int b = a + 1; b < nc; ++b)
                        outs_pair[a * nc + b] = ws + cross_off[cats[a]][cats[b]];
                }
                // by-product of the column-owner kernel below: the diagonal of the sparse block's
                // own sandwich (TABMAT_B200_SPARSE_DIAG=0: the CSR kernel adds it itself).  The
                // CSR kernel then only adds the strictly lower triangle - L2 REDs, few threads
                // busy - and can run on a side stream under the shared-memory-bound pairs kernel
                // (TABMAT_B200_INDEX_OVERLAP=1; the pairs kernel is launched first so that its one
                // CTA per SM is resident before the CSR kernel's CTAs fill the rest).
                static const bool diag_off = getenv("TABMAT_B200_SPARSE_DIAG") &&
                                             atoi(getenv("TABMAT_B200_SPARSE_DIAG")) == 0;
                static const bool overlap_on = getenv("TABMAT_B200_INDEX_OVERLAP") &&
                                               atoi(getenv("TABMAT_B200_INDEX_OVERLAP")) == 1;
                F* sdiag = nullptr;
                if (want_cs && blk[sparse_idx].csc_row_blocks > 1 && !diag_off) {
                    const tm_block_desc& S = blk[sparse_idx];
                    sdiag = ws + self_off[sparse_idx];
                    TM_CUDA(cudaMemsetAsync(sdiag, 0, sizeof(F) * (size_t)(S.ncols * S.ncols), st));
                    if (overlap_on && g_split_sched == 3 && side_stream(0)) {
                        sparse_side = side_stream(0);
                        TM_CUDA(cudaEventRecord(side_event(2), st));
                        TM_CUDA(cudaStreamWaitEvent(sparse_side, side_event(2), 0));
                    }
                }
                rc = index_cat_pairs<F>(pairs_direct_off ? rec.p : nullptr, dd, cc, dfc, n, nc, Kc,
                                        runc, outs_self, outs_pair, st);
                if (rc) return rc;
                cats_fused = true;
                if (sparse_side) {
                    const tm_block_desc& S = blk[sparse_idx];
                    rc = sparse_sandwich_ex<F>(static_cast<const F*>(S.data), S.csr_indices,
                                               S.csr_indptr, S.csr_row, n, S.ncols, S.nnz, d, rows,
                                               n_rows, (const int32_t*)nullptr, (int64_t)0, sdiag,
                                               sparse_side, true);
                    if (rc) return rc;
                }
                if (want_cs) {
                    const tm_block_desc& S = blk[sparse_idx];
                    F* outs[8];
                    for (int a = 0; a < nc; ++a) {
                        const int lo = cats[a] < sparse_idx ? cats[a] : sparse_idx;
                        const int hi = cats[a] < sparse_idx ? sparse_idx : cats[a];
                        outs[a] = ws + cross_off[lo][hi];
                    }
                    rc = index_cat_sparse<F>(use_packed ? nullptr : rec.p, dd,
                                             use_packed ? S.csc_cat_codes : nullptr, nc, Kc, runp,
                                             static_cast<const F*>(S.csc_data), S.csc_indices,
                                             S.csc_indptr, S.ncols,
                                             (int)(S.csc_row_blocks > 1 ? S.csc_row_blocks : 1),
                                             outs, sdiag, S.ncols + 1, st);
                    if (rc) return rc;
                    cat_sparse_fused = true;
                    sparse_diag_done = sdiag != nullptr;
                }
            }
        }
        for (int i = 0; i < nb; ++i) {
            const tm_block_desc& bi = blk[i];
            F* so = ws + self_off[i];
            int rc = 0;
            if (bi.kind == KIND_DENSE && dense_self_done)
                rc = 0;
            else if (bi.kind == KIND_CAT && cats_fused)
                rc = 0;
            else if (bi.kind == KIND_DENSE)
                rc = dense_sandwich(tag, static_cast<const F*>(bi.data), n, bi.ncols, bi.c_order, d,
                                    rows, n_rows, (const int32_t*)nullptr, (int64_t)0, so, stream);
            else if (bi.kind == KIND_SPARSE && sparse_diag_done && i == sparse_idx && sparse_side)
                rc = 0;   // already running on the side stream
            else if (bi.kind == KIND_SPARSE && sparse_diag_done && i == sparse_idx)
                rc = sparse_sandwich_ex<F>(static_cast<const F*>(bi.data), bi.csr_indices,
                                           bi.csr_indptr, bi.csr_row, n, bi.ncols, bi.nnz, d, rows,
                                           n_rows, (const int32_t*)nullptr, (int64_t)0, so, st, true);
            else if (bi.kind == KIND_SPARSE)
                rc = sparse_sandwich(tag, static_cast<const F*>(bi.data), bi.csr_indices,
                                     bi.csr_indptr, bi.csr_row, n, bi.ncols, bi.nnz, d, rows,
                                     n_rows, (const int32_t*)nullptr, (int64_t)0, so, stream);
            else
                rc = cat_sandwich(tag, static_cast<const int32_t*>(bi.data), n, d, rows, n_rows,
                                  bi.ncols, (int)bi.drop_first, so, stream);
            if (rc) return rc;
            for (int j = i + 1; j < nb; ++j) {
                const tm_block_desc& bj = blk[j];
                F* co = ws + cross_off[i][j];
                const bool has_dense = bi.kind == KIND_DENSE || bj.kind == KIND_DENSE;
                if (fuse && has_dense) continue;
                if (cats_fused && bi.kind == KIND_CAT && bj.kind == KIND_CAT) continue;
                if (cat_sparse_fused && ((bi.kind == KIND_CAT && bj.kind == KIND_SPARSE) ||
                                         (bi.kind == KIND_SPARSE && bj.kind == KIND_CAT)))
                    continue;
                // normalise to (a, b) = the kernel's native (rows, cols) orientation
                const tm_block_desc& a = cross_rows_are_j(bi, bj) ? bj : bi;
                const tm_block_desc& b = cross_rows_are_j(bi, bj) ? bi : bj;
                if (a.kind == KIND_SPARSE && b.kind == KIND_DENSE)
                    rc = csr_dense_sandwich(tag, static_cast<const F*>(a.data), a.csr_indices,
                                            a.csr_indptr, n, a.ncols,
                                            static_cast<const F*>(b.data), b.ncols,
                                            (int)b.c_order, d, rows, n_rows,
                                            (const int32_t*)nullptr, (int64_t)0,
                                            (const int32_t*)nullptr, (int64_t)0, co, stream);
                else if (a.kind == KIND_CAT && b.kind == KIND_DENSE)
                    rc = cat_dense_sandwich(tag, static_cast<const int32_t*>(a.data), n, a.ncols,
                                            (int)a.drop_first, d, static_cast<const F*>(b.data),
                                            b.ncols, (int)b.c_order, rows, n_rows,
                                            (const int32_t*)nullptr, (int64_t)0, co, stream);
                else if (a.kind == KIND_CAT && b.kind == KIND_SPARSE)
                    rc = cat_sparse_sandwich(tag, static_cast<const int32_t*>(a.data), n, a.ncols,
                                             (int)a.drop_first, d, static_cast<const F*>(b.data),
                                             b.csr_indices, b.csr_indptr, b.csr_row, b.ncols,
                                             b.nnz, rows, n_rows, (const int32_t*)nullptr,
                                             (int64_t)0, co, stream);
                else if (a.kind == KIND_CAT && b.kind == KIND_CAT)
                    rc = cat_cat_sandwich(tag, static_cast<const int32_t*>(a.data),
                                          static_cast<const int32_t*>(b.data), n, a.ncols, b.ncols,
                                          (int)a.drop_first, (int)b.drop_first, d, rows, n_rows,
                                          co, stream);
                else
                    return fail("tm_split_sandwich: unsupported block pair (two dense or two "
                                "sparse blocks must be merged first, split_matrix.py:85-141)");
                if (rc) return rc;
            }
        }
        if (sparse_side) {
            TM_CUDA(cudaEventRecord(side_event(3), sparse_side));
            TM_CUDA(cudaStreamWaitEvent(st, side_event(3), 0));
        }
        pass_mark(PASS_INDEX, 1, st);
        return 0;
    };

    // ---- schedule ---------------------------------------------------------------------------
    // T = tensor pass (HBM / tensor pipe), I = index pass (shared-memory atomics + gathers),
    // S = scatter pass (L2 atomic units).  TABMAT_B200_SCHED:
    //   0  side: T   | main: I, S        1  side: T | main: S, I
    //   2  side1: T  | side2: I | main: S (everything concurrent)
    //   3  main: T, I, S (serial)        4  main: T, then side: I | main: S
    //   5  main: I, then side: T | main: S
    // TABMAT_B200_SIDE_STREAM=0 forces 3.
    cudaStream_t main_st = as_stream(stream);
    if (v && dense_vec && dense_idx >= 0 && n_dense == 1 && !tc_ok && part != 1) {
        const tm_block_desc& D = blk[dense_idx];
        int rcv = dense_rmatvec(tag, static_cast<const F*>(D.data), n, D.ncols, (int)D.c_order, v,
                                rows, n_rows, (const int32_t*)nullptr, (int64_t)0, dense_vec, stream);
        if (rcv) return rcv;
    }
    cudaStream_t s1 = side_stream(0), s2 = side_stream(1);
    int sched = g_split_sched;
    if (!s1 || !s2) sched = 3;
    if (part == 1) {  // only the blocks without the dense operand (tm_split_sandwich_blocks_part)
        int rc1 = index_pass(main_st);
        if (g_profile && rc1 == 0) ++g_prof_calls;
        return rc1;
    }
    if (part == 2) {  // the rest
        // with SMs reserved for a collective that a row-sharded caller started after part 1
        // (tm_set_sm_reserve), the gather form goes first: it can share the GPU, the persistent
        // tcgen05 kernel cannot
        int rc2 = 0;
        if (sparse_by_gather && g_sm_reserve > 0) {
            rc2 = scatter_pass(main_st);
            if (rc2 == 0) rc2 = tensor_pass(main_st, false);
        } else {
            rc2 = tensor_pass(main_st, false);
            if (rc2 == 0) rc2 = scatter_pass(main_st);
        }
        return rc2;
    }
    if (!have_tensor && (sched == 0 || sched == 1)) sched = 3;
    auto fork_to = [&](cudaStream_t st, int ev) -> int {
        TM_CUDA(cudaEventRecord(side_event(ev), main_st));
        TM_CUDA(cudaStreamWaitEvent(st, side_event(ev), 0));
        return 0;
    };
    auto join_from = [&](cudaStream_t st, int ev) -> int {
        TM_CUDA(cudaEventRecord(side_event(ev), st));
        TM_CUDA(cudaStreamWaitEvent(main_st, side_event(ev), 0));
        return 0;
    };
    int rc = 0;
#define TM_RC(x)          \
    do {                  \
        rc = (x);         \
        if (rc) return rc; \
    } while (0)
    switch (sched) {
        case 0:
        case 1:
            TM_RC(fork_to(s1, 0));
            TM_RC(tensor_pass(s1, true));
            if (sched == 0) {
                TM_RC(index_pass(main_st));
                TM_RC(scatter_pass(main_st));
            } else {
                TM_RC(scatter_pass(main_st));
                TM_RC(index_pass(main_st));
            }
            TM_RC(join_from(s1, 1));
            break;
        case 2:
            TM_RC(fork_to(s1, 0));
            TM_RC(fork_to(s2, 2));
            TM_RC(tensor_pass(s1, true));
            TM_RC(scatter_pass(main_st));
            TM_RC(index_pass(s2));
            TM_RC(join_from(s1, 1));
            TM_RC(join_from(s2, 3));
            break;
        case 4:
            TM_RC(tensor_pass(main_st, false));
            TM_RC(fork_to(s2, 2));
            TM_RC(scatter_pass(main_st));
            TM_RC(index_pass(s2));
            TM_RC(join_from(s2, 3));
            break;
        case 5:
            TM_RC(index_pass(main_st));
            TM_RC(fork_to(s1, 0));
            TM_RC(tensor_pass(s1, true));
            TM_RC(scatter_pass(main_st));
            TM_RC(join_from(s1, 1));
            break;
        default:
            TM_RC(tensor_pass(main_st, false));
            TM_RC(index_pass(main_st));
            TM_RC(scatter_pass(main_st));
            break;
    }
#undef TM_RC
    if (g_profile) ++g_prof_calls;
    return 0;
}

// ---- placement of all blocks in ONE launch -------------------------------------------------
// A job = one stored block (na x nb at `src`, destination rows ri / columns ci; diag: the
// vector of a categorical self block).  32 x 32 tiles of all jobs are numbered consecutively;
// a CTA finds its job by a linear scan of the (at most 64) tile prefixes.  The per-block
// launches this replaces cost ~5 us each (28 of them at the benchmark shape = half of the
// placement time, and 4 % of a step at 8 GPUs).
constexpr int ASM_MAX_JOBS = 64;
struct AsmJobs {
    const void* src[ASM_MAX_JOBS];
    const int64_t* ri[ASM_MAX_JOBS];
    const int64_t* ci[ASM_MAX_JOBS];
    int na[ASM_MAX_JOBS];
    int nb[ASM_MAX_JOBS];
    int tile0[ASM_MAX_JOBS + 1];   // first tile of the job
    unsigned char kind[ASM_MAX_JOBS];  // 0 block, 1 block + mirror, 2 diagonal
    int n_jobs;
};

// row0 / row1: only destination rows in [row0, row1) are written, at out + (row - row0) * ld —
// a row band of the result (each rank of a row-sharded job places and copies out its own band)
template <typename F>
__global__ void __launch_bounds__(256)
k_assemble_all(const AsmJobs jobs, double* __restrict__ out, int64_t ld, int64_t row0,
               int64_t row1) {
    __shared__ F tile[32][33];
    const int t = blockIdx.x;
    int q = 0;
    while (q + 1 < jobs.n_jobs && t >= jobs.tile0[q + 1]) ++q;
    const int na = jobs.na[q], nb = jobs.nb[q];
    const int tiles_x = (nb + 31) / 32;
    const int local = t - jobs.tile0[q];
    const int64_t a0 = (int64_t)(local / tiles_x) * 32, b0 = (int64_t)(local % tiles_x) * 32;
    const int64_t* __restrict__ ri = jobs.ri[q];
    const int64_t* __restrict__ ci = jobs.ci[q];
    const F* __restrict__ src = static_cast<const F*>(jobs.src[q]);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int kind = jobs.kind[q];
    if (kind == 2) {  // diagonal block: na == nb, src = the diagonal
        for (int i = ty; i < 32; i += 8) {
            const int64_t a = a0 + i, b = b0 + tx;
            if (a < na && b < nb && ri[a] >= row0 && ri[a] < row1 && ri[b] >= 0)
                out[(ri[a] - row0) * ld + ri[b]] = a == b ? (double)src[a] : 0.0;
        }
        return;
    }
    // a negative destination = the column is not in the caller's `cols` selection (row0 >= 0)
    for (int i = ty; i < 32; i += 8) {
        const int64_t a = a0 + i, b = b0 + tx;
        if (a < na && b < nb) {
            const F v = src[a * nb + b];
            tile[i][tx] = v;
            if (ri[a] >= row0 && ri[a] < row1 && ci[b] >= 0)
                out[(ri[a] - row0) * ld + ci[b]] = (double)v;
        }
    }
    if (kind != 1) return;
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int64_t b = b0 + i, a = a0 + tx;
        if (a < na && b < nb && ri[a] >= 0 && ci[b] >= row0 && ci[b] < row1)
            out[(ci[b] - row0) * ld + ri[a]] = (double)tile[tx][i];
    }
}

template <typename F>
int split_assemble(const tm_block_desc* blk, int nb, const F* ws, double* out, int64_t ld,
                   tm_stream_t stream, int part = 0, int64_t row0 = 0,
                   int64_t row1 = INT64_MAX) {
    F* tag = nullptr;
    int64_t off = 0;
    // part 1: blocks without a dense operand; part 2: blocks with one; 0: all
    auto wanted = [part](bool has_dense) { return part == 0 || (part == 2) == has_dense; };
    static const bool one_launch =
        !(getenv("TABMAT_B200_ASSEMBLE_FUSED") && atoi(getenv("TABMAT_B200_ASSEMBLE_FUSED")) == 0);
    bool all_indexed = true;
    for (int i = 0; i < nb; ++i) all_indexed &= blk[i].col_index != nullptr;
    const bool band = row0 > 0 || row1 != INT64_MAX;
    if (band && !(all_indexed && nb * (nb + 1) / 2 <= ASM_MAX_JOBS))
        return fail("tm_split_sandwich_assemble_band: needs col_index on every block, <= 10 blocks");
    if ((one_launch || band) && all_indexed && nb * (nb + 1) / 2 <= ASM_MAX_JOBS) {
        AsmJobs jobs;
        memset(&jobs, 0, sizeof(jobs));
        int nj = 0;
        int64_t tiles = 0;
        auto add = [&](const F* src, const tm_block_desc& a, const tm_block_desc& b, int kind) {
            if (a.ncols <= 0 || b.ncols <= 0) return;
            jobs.src[nj] = src;
            jobs.ri[nj] = a.col_index;
            jobs.ci[nj] = b.col_index;
            jobs.na[nj] = (int)a.ncols;
            jobs.nb[nj] = (int)b.ncols;
            jobs.kind[nj] = (unsigned char)kind;
            jobs.tile0[nj] = (int)tiles;
            tiles += ((a.ncols + 31) / 32) * ((b.ncols + 31) / 32);
            ++nj;
        };
        for (int i = 0; i < nb; ++i) {
            const tm_block_desc& bi = blk[i];
            if (wanted(bi.kind == KIND_DENSE)) add(ws + off, bi, bi, bi.kind == KIND_CAT ? 2 : 0);
            off += ws_pad(self_elems(bi));
            for (int j = i + 1; j < nb; ++j) {
                const tm_block_desc& bj = blk[j];
                const tm_block_desc& a = cross_rows_are_j(bi, bj) ? bj : bi;
                const tm_block_desc& b = cross_rows_are_j(bi, bj) ? bi : bj;
                if (wanted(bi.kind == KIND_DENSE || bj.kind == KIND_DENSE)) add(ws + off, a, b, 1);
                off += ws_pad(bi.ncols * bj.ncols);
            }
        }
        jobs.tile0[nj] = (int)tiles;
        jobs.n_jobs = nj;
        if (nj == 0 || tiles == 0) return 0;
        if (tiles < (int64_t)kMaxGridX) {
            k_assemble_all<F><<<(unsigned)tiles, 256, 0, as_stream(stream)>>>(jobs, out, ld, row0,
                                                                              row1);
            TM_LAUNCHED();
            return 0;
        }
        if (band) return fail("tm_split_sandwich_assemble_band: result too large");
        off = 0;  // absurdly large: fall through to the per-block launches
    }
    for (int i = 0; i < nb; ++i) {
        const tm_block_desc& bi = blk[i];
        int rc = 0;
        if (!wanted(bi.kind == KIND_DENSE))
            rc = 0;
        else if (bi.kind == KIND_CAT)
            rc = scatter_diag(tag, ws + off, bi.ncols, bi.col_index, out, ld, stream);
        else
            rc = scatter_block(tag, ws + off, bi.ncols, bi.ncols, bi.col_index, bi.col_index, out,
                               ld, 0, stream);
        if (rc) return rc;
        off += ws_pad(self_elems(bi));
        for (int j = i + 1; j < nb; ++j) {
            const tm_block_desc& bj = blk[j];
            const tm_block_desc& a = cross_rows_are_j(bi, bj) ? bj : bi;
            const tm_block_desc& b = cross_rows_are_j(bi, bj) ? bi : bj;
            if (wanted(bi.kind == KIND_DENSE || bj.kind == KIND_DENSE)) {
                rc = scatter_block(tag, ws + off, a.ncols, b.ncols, a.col_index, b.col_index, out,
                                   ld, 1, stream);
                if (rc) return rc;
            }
            off += ws_pad(bi.ncols * bj.ncols);
        }
    }
    return 0;
}

}  // namespace tmb

extern "C" {

int64_t tm_sizeof_block_desc(void) { return (int64_t)sizeof(tm_block_desc); }
int tm_split_last_plan(void) { return tmb::g_last_plan; }

void tm_split_profile_enable(int on) {
    tmb::g_profile = on != 0;
    tmb::g_prof_calls = 0;
    for (int c = 0; c < tmb::PROF_RING; ++c)
        for (int p = 0; p < tmb::PASS_COUNT; ++p) tmb::g_pass_used[c][p] = false;
}
int tm_split_profile_read(float* ms) {
    const int calls = tmb::g_prof_calls < tmb::PROF_RING ? tmb::g_prof_calls : tmb::PROF_RING;
    for (int p = 0; p < tmb::PASS_COUNT; ++p) {
        double sum = 0;
        int cnt = 0;
        for (int c = 0; c < calls; ++c) {
            if (!tmb::g_pass_used[c][p]) continue;
            float t = 0.f;
            if (cudaEventSynchronize(tmb::g_pass_ev[c][p][1]) != cudaSuccess) return 1;
            cudaEventElapsedTime(&t, tmb::g_pass_ev[c][p][0], tmb::g_pass_ev[c][p][1]);
            sum += t;
            ++cnt;
        }
        ms[p] = cnt ? (float)(sum / cnt) : -1.f;
    }
    return 0;
}

int64_t tm_split_workspace_elems(const tm_block_desc* blocks, int n_blocks) {
    int64_t off = 0;
    for (int i = 0; i < n_blocks; ++i) {
        off += tmb::ws_pad(tmb::self_elems(blocks[i]));
        for (int j = i + 1; j < n_blocks; ++j)
            off += tmb::ws_pad(blocks[i].ncols * blocks[j].ncols);
    }
    return off;
}

int64_t tm_split_workspace_head_elems(const tm_block_desc* blocks, int n_blocks) {
    if (n_blocks <= 0) return 0;
    int64_t off = tmb::ws_pad(tmb::self_elems(blocks[0]));
    for (int j = 1; j < n_blocks; ++j) off += tmb::ws_pad(blocks[0].ncols * blocks[j].ncols);
    return off;
}

int tm_split_sandwich_blocks_f32(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                 const float* d, const int32_t* rows, int64_t n_rows,
                                 float* workspace, tm_stream_t stream) {
    return tmb::split_blocks<float>(blocks, n_blocks, n, d, rows, n_rows, workspace, stream);
}
int tm_split_sandwich_blocks_f64(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                 const double* d, const int32_t* rows, int64_t n_rows,
                                 double* workspace, tm_stream_t stream) {
    return tmb::split_blocks<double>(blocks, n_blocks, n, d, rows, n_rows, workspace, stream);
}
int tm_split_sandwich_rmatvec_blocks_f32(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                         const float* d, const float* v, const int32_t* rows,
                                         int64_t n_rows, float* workspace, float* dense_vec,
                                         tm_stream_t stream) {
    if (!v || !dense_vec) return tmb::fail("tm_split_sandwich_rmatvec_blocks: v and dense_vec are required");
    return tmb::split_blocks<float>(blocks, n_blocks, n, d, rows, n_rows, workspace, stream, 0, v,
                                    dense_vec);
}
int tm_split_sandwich_rmatvec_blocks_f64(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                         const double* d, const double* v, const int32_t* rows,
                                         int64_t n_rows, double* workspace, double* dense_vec,
                                         tm_stream_t stream) {
    if (!v || !dense_vec) return tmb::fail("tm_split_sandwich_rmatvec_blocks: v and dense_vec are required");
    return tmb::split_blocks<double>(blocks, n_blocks, n, d, rows, n_rows, workspace, stream, 0, v,
                                     dense_vec);
}
int tm_split_sandwich_blocks_part_f32(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                      const float* d, const int32_t* rows, int64_t n_rows,
                                      float* workspace, int part, tm_stream_t stream) {
    return tmb::split_blocks<float>(blocks, n_blocks, n, d, rows, n_rows, workspace, stream, part);
}
int tm_split_sandwich_blocks_part_f64(const tm_block_desc* blocks, int n_blocks, int64_t n,
                                      const double* d, const int32_t* rows, int64_t n_rows,
                                      double* workspace, int part, tm_stream_t stream) {
    return tmb::split_blocks<double>(blocks, n_blocks, n, d, rows, n_rows, workspace, stream, part);
}
int tm_split_sandwich_assemble_part_f32(const tm_block_desc* blocks, int n_blocks,
                                        const float* workspace, double* out, int64_t ld, int part,
                                        tm_stream_t stream) {
    return tmb::split_assemble<float>(blocks, n_blocks, workspace, out, ld, stream, part);
}
int tm_split_sandwich_assemble_part_f64(const tm_block_desc* blocks, int n_blocks,
                                        const double* workspace, double* out, int64_t ld, int part,
                                        tm_stream_t stream) {
    return tmb::split_assemble<double>(blocks, n_blocks, workspace, out, ld, stream, part);
}
int tm_split_sandwich_assemble_band_f32(const tm_block_desc* blocks, int n_blocks,
                                        const float* workspace, double* out_band, int64_t ld,
                                        int64_t row0, int64_t row1, tm_stream_t stream) {
    if (row0 < 0 || row1 < row0) return tmb::fail("tm_split_sandwich_assemble_band: bad row range");
    return tmb::split_assemble<float>(blocks, n_blocks, workspace, out_band, ld, stream, 0, row0,
                                      row1);
}
int tm_split_sandwich_assemble_band_f64(const tm_block_desc* blocks, int n_blocks,
                                        const double* workspace, double* out_band, int64_t ld,
                                        int64_t row0, int64_t row1, tm_stream_t stream) {
    if (row0 < 0 || row1 < row0) return tmb::fail("tm_split_sandwich_assemble_band: bad row range");
    return tmb::split_assemble<double>(blocks, n_blocks, workspace, out_band, ld, stream, 0, row0,
                                       row1);
}
int tm_split_sandwich_assemble_part_band_f32(const tm_block_desc* blocks, int n_blocks,
                                             const float* workspace, double* out_band, int64_t ld,
                                             int part, int64_t row0, int64_t row1,
                                             tm_stream_t stream) {
    if (row0 < 0 || row1 < row0) return tmb::fail("tm_split_sandwich_assemble_part_band: bad row range");
    return tmb::split_assemble<float>(blocks, n_blocks, workspace, out_band, ld, stream, part, row0,
                                      row1);
}
int tm_split_sandwich_assemble_part_band_f64(const tm_block_desc* blocks, int n_blocks,
                                             const double* workspace, double* out_band, int64_t ld,
                                             int part, int64_t row0, int64_t row1,
                                             tm_stream_t stream) {
    if (row0 < 0 || row1 < row0) return tmb::fail("tm_split_sandwich_assemble_part_band: bad row range");
    return tmb::split_assemble<double>(blocks, n_blocks, workspace, out_band, ld, stream, part, row0,
                                       row1);
}
int tm_memcpy2d_to_host(void* dst_host, int64_t dst_pitch, const void* src_dev, int64_t src_pitch,
                        int64_t width_bytes, int64_t height, tm_stream_t stream) {
    if (width_bytes <= 0 || height <= 0) return 0;
    TM_CUDA(cudaMemcpy2DAsync(dst_host, (size_t)dst_pitch, src_dev, (size_t)src_pitch,
                              (size_t)width_bytes, (size_t)height, cudaMemcpyDeviceToHost,
                              tmb::as_stream(stream)));
    return 0;
}
int tm_split_sandwich_assemble_f32(const tm_block_desc* blocks, int n_blocks,
                                   const float* workspace, double* out, int64_t ld,
                                   tm_stream_t stream) {
    return tmb::split_assemble<float>(blocks, n_blocks, workspace, out, ld, stream);
}
int tm_split_sandwich_assemble_f64(const tm_block_desc* blocks, int n_blocks,
                                   const double* workspace, double* out, int64_t ld,
                                   tm_stream_t stream) {
    return tmb::split_assemble<double>(blocks, n_blocks, workspace, out, ld, stream);
}

}  // extern "C"
