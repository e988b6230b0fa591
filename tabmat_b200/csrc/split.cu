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
#include <cstdlib>
#include <vector>

#include "tm_common.cuh"

namespace tmb {

enum { KIND_DENSE = 0, KIND_SPARSE = 1, KIND_CAT = 2 };

static inline int64_t self_elems(const tm_block_desc& b) {
    return b.kind == KIND_CAT ? b.ncols : b.ncols * b.ncols;
}

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

// one non-blocking side stream + two events per thread (created lazily, never destroyed);
// TABMAT_B200_SIDE_STREAM=0 keeps everything on the caller's stream
static cudaStream_t side_stream() {
    static thread_local cudaStream_t st = nullptr;
    static thread_local bool tried = false;
    if (!tried) {
        tried = true;
        const char* e = getenv("TABMAT_B200_SIDE_STREAM");
        if (!(e && atoi(e) == 0)) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    }
    return st;
}
static cudaEvent_t side_event(int i) {
    static thread_local cudaEvent_t ev[2] = {nullptr, nullptr};
    if (!ev[i]) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    return ev[i];
}

// ---- optional pass-level timing (tm_split_profile_*): CUDA events on the stream each pass
// is launched on; bench.py reads them after synchronising ---------------------------------
enum { PASS_TENSOR = 0, PASS_SCATTER = 1, PASS_INDEX = 2, PASS_COUNT = 3 };
constexpr int PROF_RING = 64;  // calls kept since tm_split_profile_enable(1)
static bool g_profile = false;
static int g_prof_calls = 0;
static cudaEvent_t g_pass_ev[PROF_RING][PASS_COUNT][2];
static bool g_pass_used[PROF_RING][PASS_COUNT];
static void pass_mark(int pass, int which, cudaStream_t st) {
    if (!g_profile) return;
    const int slot = g_prof_calls % PROF_RING;
    if (!g_pass_ev[slot][pass][which]) cudaEventCreate(&g_pass_ev[slot][pass][which]);
    cudaEventRecord(g_pass_ev[slot][pass][which], st);
    if (which == 1) g_pass_used[slot][pass] = true;
}

static inline int64_t n_rows_or_all(const int32_t* rows, int64_t n_rows, int64_t n) {
    return rows ? n_rows : n;
}

template <typename F>
int split_blocks(const tm_block_desc* blk, int nb, int64_t n, const F* d, const int32_t* rows,
                 int64_t n_rows, F* ws, tm_stream_t stream) {
    F* tag = nullptr;
    if (nb <= 0) return 0;
    if (nb > 64) return fail("tm_split_sandwich: more than 64 blocks");
    // offsets
    std::vector<int64_t> self_off(nb);
    std::vector<std::vector<int64_t>> cross_off(nb, std::vector<int64_t>(nb, -1));
    int64_t off = 0;
    for (int i = 0; i < nb; ++i) {
        self_off[i] = off;
        off += self_elems(blk[i]);
        for (int j = i + 1; j < nb; ++j) {
            cross_off[i][j] = off;
            off += blk[i].ncols * blk[j].ncols;
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
    // HBM / tensor bound work (the tcgen05 pass and the sorted-gather passes) goes to a side
    // stream so that it overlaps with the L2-atomic-bound scatter passes on the caller's stream
    std::vector<char> on_tensor(nb, 0);
    bool dense_self_done = false;
    bool side_used = false;
    if (fuse && sizeof(F) == 4) {
        const tm_block_desc& D = blk[dense_idx];
        static const bool onehot_off =
            getenv("TABMAT_B200_ONEHOT") && atoi(getenv("TABMAT_B200_ONEHOT")) == 0;
        static const bool gather_off =
            getenv("TABMAT_B200_GATHER") && atoi(getenv("TABMAT_B200_GATHER")) == 0;
        const bool tc_ok = g_dense_f32_mode != 1 && n_rows_or_all(rows, n_rows, n) > 0 &&
                           dense_tc_eligible(n, D.ncols, 1, D.data);
        bool any_gather = false;
        for (int i = 0; i < nb; ++i)
            any_gather |= !gather_off && blk[i].kind == KIND_CAT && blk[i].cat_perm != nullptr;
        if (tc_ok || any_gather) {
            cudaStream_t st = side_stream() ? side_stream() : as_stream(stream);
            side_used = st != as_stream(stream);
            if (side_used) {
                TM_CUDA(cudaEventRecord(side_event(0), as_stream(stream)));
                TM_CUDA(cudaStreamWaitEvent(st, side_event(0), 0));
            }
            pass_mark(PASS_TENSOR, 0, st);
            Scratch dm(rows ? sizeof(float) * (size_t)n : 0, st);
            if (dm.err != cudaSuccess) return fail_cuda(dm.err, "scratch");
            const float* dd = reinterpret_cast<const float*>(d);
            if (rows) {
                int rc = masked_weights<float>(dd, n, rows, n_rows, dm.as<float>(), st);
                if (rc) return rc;
                dd = dm.as<float>();
            }
            if (tc_ok) {
                TcOneHot oh;
                oh.ncat = 0;
                int64_t slots = 0;
                int which[8];
                for (int i = 0; i < nb && oh.ncat < 8 && D.ncols <= 128 && !onehot_off; ++i) {
                    if (blk[i].kind != KIND_CAT || blk[i].ncols <= 0 || blk[i].ncols > 256) continue;
                    if (slots + blk[i].ncols > TC_ONEHOT_MAX_SLOTS) continue;
                    which[oh.ncat] = i;
                    oh.codes[oh.ncat] = static_cast<const int32_t*>(blk[i].data);
                    oh.K[oh.ncat] = (int)blk[i].ncols;
                    oh.drop_first[oh.ncat] = blk[i].drop_first;
                    slots += blk[i].ncols;
                    ++oh.ncat;
                }
                Scratch tmp(sizeof(float) * (size_t)(slots > 0 ? slots : 1) * (size_t)D.ncols, st);
                if (tmp.err != cudaSuccess) return fail_cuda(tmp.err, "scratch");
                oh.out = tmp.as<float>();
                int rc = dense_sandwich_tc_f32(static_cast<const float*>(D.data), n, D.ncols, 1, dd,
                                               reinterpret_cast<float*>(ws + self_off[dense_idx]),
                                               st, oh.ncat ? &oh : nullptr, /*share_sm=*/side_used);
                if (rc) return rc;
                dense_self_done = true;
                int64_t o = 0;
                for (int c = 0; c < oh.ncat; ++c) {
                    int i = which[c];
                    int a = i < dense_idx ? i : dense_idx, b = i < dense_idx ? dense_idx : i;
                    TM_CUDA(cudaMemcpyAsync(ws + cross_off[a][b], tmp.as<float>() + o * D.ncols,
                                            sizeof(float) * (size_t)(blk[i].ncols * D.ncols),
                                            cudaMemcpyDeviceToDevice, st));
                    o += blk[i].ncols;
                    on_tensor[i] = 1;
                }
            }
            // categorical blocks with a sorted row permutation: HBM-bound gather
            for (int i = 0; i < nb && !gather_off; ++i) {
                if (blk[i].kind != KIND_CAT || on_tensor[i] || !blk[i].cat_perm) continue;
                int a = i < dense_idx ? i : dense_idx, b = i < dense_idx ? dense_idx : i;
                int rc = cat_dense_gather_f32(static_cast<const float*>(D.data), D.ncols, dd,
                                              blk[i].cat_perm, blk[i].cat_segptr, blk[i].ncols,
                                              blk[i].cat_nvalid,
                                              reinterpret_cast<float*>(ws + cross_off[a][b]), st);
                if (rc) return rc;
                on_tensor[i] = 1;
            }
            pass_mark(PASS_TENSOR, 1, st);
        }
    }
    auto scatter_pass = [&]() -> int {
    if (fuse) {
            const tm_block_desc& D = blk[dense_idx];
            const int32_t* codes[8];
            int64_t K[8];
            int32_t df[8];
            F* outs[8];
            int c = 0;
            for (int i = 0; i < nb; ++i) {
                if (blk[i].kind != KIND_CAT || on_tensor[i]) continue;
                codes[c] = static_cast<const int32_t*>(blk[i].data);
                K[c] = blk[i].ncols;
                df[c] = blk[i].drop_first;
                int a = i < dense_idx ? i : dense_idx, b = i < dense_idx ? dense_idx : i;
                outs[c] = ws + cross_off[a][b];
                ++c;
            }
            const F* sdata = nullptr;
            const int32_t *sind = nullptr, *sptr = nullptr;
            int64_t ps = 0;
            F* out_s = nullptr;
            if (sparse_idx >= 0 && blk[sparse_idx].nnz > 0) {
                const tm_block_desc& S = blk[sparse_idx];
                sdata = static_cast<const F*>(S.data);
                sind = S.csr_indices;
                sptr = S.csr_indptr;
                ps = S.ncols;
                int a = sparse_idx < dense_idx ? sparse_idx : dense_idx;
                int b = sparse_idx < dense_idx ? dense_idx : sparse_idx;
                out_s = ws + cross_off[a][b];
            } else if (sparse_idx >= 0) {
                int a = sparse_idx < dense_idx ? sparse_idx : dense_idx;
                int b = sparse_idx < dense_idx ? dense_idx : sparse_idx;
                TM_CUDA(cudaMemsetAsync(ws + cross_off[a][b], 0,
                                        sizeof(F) * (size_t)(blk[a].ncols * blk[b].ncols),
                                        as_stream(stream)));
            }
            if (c > 0 || out_s) {
                pass_mark(PASS_SCATTER, 0, as_stream(stream));
                int rc = dense_cross_fused<F>(static_cast<const F*>(D.data), n, D.ncols, d, rows,
                                              n_rows, c, codes, K, df, outs, sdata, sind, sptr, ps,
                                              out_s, /*runs=*/1, as_stream(stream));
                if (rc) return rc;
                pass_mark(PASS_SCATTER, 1, as_stream(stream));
            }
        }
        return 0;
    };
    // order of the two passes on the caller's stream (the index pass first lets the lighter
    // kernels share the SMs with the tensor pass; TABMAT_B200_SCATTER_FIRST=1 restores the other)
    static const bool scatter_first =
        getenv("TABMAT_B200_SCATTER_FIRST") && atoi(getenv("TABMAT_B200_SCATTER_FIRST")) == 1;
    if (scatter_first) {
        int rc = scatter_pass();
        if (rc) return rc;
    }
    pass_mark(PASS_INDEX, 0, as_stream(stream));

    // ---- fused index blocks (split_index.cu): all categorical self / pair blocks in one pass
    // over 32-byte row records, categorical x sparse for all categorical blocks from the CSC
    // copy without global atomics ----------------------------------------------------------
    bool cats_fused = false, cat_sparse_fused = false;
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
        int32_t dfc[8], runc[8];
        bool ok = nc >= 1 && nc <= 7 && n_rows_or_all(rows, n_rows, n) > 0 && n > 0;
        if (ok) {
            for (int a = 0; a < nc; ++a) {
                const tm_block_desc& b = blk[cats[a]];
                Kc[a] = b.ncols;
                cc[a] = static_cast<const int32_t*>(b.data);
                dfc[a] = b.drop_first;
                runc[a] = (b.flags & TM_BLOCK_FLAG_RUNS) ? 1 : 0;
            }
            ok = index_fused_eligible<F>(nc, Kc);
        }
        Scratch rec(ok ? index_record_bytes<F>(n) : 0, as_stream(stream));
        Scratch dmi(ok && rows ? sizeof(F) * (size_t)n : 0, as_stream(stream));
        if (rec.err != cudaSuccess) return fail_cuda(rec.err, "scratch");
        if (dmi.err != cudaSuccess) return fail_cuda(dmi.err, "scratch");
        if (ok) {
            const F* dd = d;
            if (rows) {
                int rc = masked_weights<F>(d, n, rows, n_rows, dmi.as<F>(), as_stream(stream));
                if (rc) return rc;
                dd = dmi.as<F>();
            }
            int rc = index_pack_records<F>(dd, n, nc, cc, dfc, rec.p, as_stream(stream));
            if (rc) return rc;
            F* outs_self[8];
            F* outs_pair[64];
            for (int a = 0; a < nc; ++a) {
                outs_self[a] = ws + self_off[cats[a]];
                for (int b = a + 1; b < nc; ++b)
                    outs_pair[a * nc + b] = ws + cross_off[cats[a]][cats[b]];
            }
            rc = index_cat_pairs<F>(rec.p, n, nc, Kc, runc, outs_self, outs_pair,
                                    as_stream(stream));
            if (rc) return rc;
            cats_fused = true;
            if (sparse_idx >= 0 && blk[sparse_idx].csc_indptr && blk[sparse_idx].csc_indices &&
                blk[sparse_idx].csc_data &&
                index_cat_sparse_fits<F>(nc, Kc, blk[sparse_idx].ncols)) {
                const tm_block_desc& S = blk[sparse_idx];
                F* outs[8];
                for (int a = 0; a < nc; ++a) {
                    const int lo = cats[a] < sparse_idx ? cats[a] : sparse_idx;
                    const int hi = cats[a] < sparse_idx ? sparse_idx : cats[a];
                    outs[a] = ws + cross_off[lo][hi];
                }
                rc = index_cat_sparse<F>(rec.p, nc, Kc, runc, static_cast<const F*>(S.csc_data),
                                         S.csc_indices, S.csc_indptr, S.ncols, outs,
                                         as_stream(stream));
                if (rc) return rc;
                cat_sparse_fused = true;
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
        else if (bi.kind == KIND_SPARSE)
            rc = sparse_sandwich(tag, static_cast<const F*>(bi.data), bi.csr_indices, bi.csr_indptr,
                                 bi.csr_row, n, bi.ncols, bi.nnz, d, rows, n_rows,
                                 (const int32_t*)nullptr, (int64_t)0, so, stream);
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
                                        a.csr_indptr, n, a.ncols, static_cast<const F*>(b.data),
                                        b.ncols, (int)b.c_order, d, rows, n_rows,
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
                                         b.csr_indices, b.csr_indptr, b.csr_row, b.ncols, b.nnz,
                                         rows, n_rows, (const int32_t*)nullptr, (int64_t)0, co,
                                         stream);
            else if (a.kind == KIND_CAT && b.kind == KIND_CAT)
                rc = cat_cat_sandwich(tag, static_cast<const int32_t*>(a.data),
                                      static_cast<const int32_t*>(b.data), n, a.ncols, b.ncols,
                                      (int)a.drop_first, (int)b.drop_first, d, rows, n_rows, co,
                                      stream);
            else
                return fail("tm_split_sandwich: unsupported block pair (two dense or two sparse "
                            "blocks must be merged first, split_matrix.py:85-141)");
            if (rc) return rc;
        }
    }
    pass_mark(PASS_INDEX, 1, as_stream(stream));
    if (!scatter_first) {
        int rc = scatter_pass();
        if (rc) return rc;
    }
    if (g_profile) ++g_prof_calls;
    if (side_used) {  // join the side stream
        TM_CUDA(cudaEventRecord(side_event(1), side_stream()));
        TM_CUDA(cudaStreamWaitEvent(as_stream(stream), side_event(1), 0));
    }
    return 0;
}

template <typename F>
int split_assemble(const tm_block_desc* blk, int nb, const F* ws, double* out, int64_t ld,
                   tm_stream_t stream) {
    F* tag = nullptr;
    int64_t off = 0;
    for (int i = 0; i < nb; ++i) {
        const tm_block_desc& bi = blk[i];
        int rc;
        if (bi.kind == KIND_CAT)
            rc = scatter_diag(tag, ws + off, bi.ncols, bi.col_index, out, ld, stream);
        else
            rc = scatter_block(tag, ws + off, bi.ncols, bi.ncols, bi.col_index, bi.col_index, out,
                               ld, 0, stream);
        if (rc) return rc;
        off += self_elems(bi);
        for (int j = i + 1; j < nb; ++j) {
            const tm_block_desc& bj = blk[j];
            const tm_block_desc& a = cross_rows_are_j(bi, bj) ? bj : bi;
            const tm_block_desc& b = cross_rows_are_j(bi, bj) ? bi : bj;
            rc = scatter_block(tag, ws + off, a.ncols, b.ncols, a.col_index, b.col_index, out, ld,
                               1, stream);
            if (rc) return rc;
            off += bi.ncols * bj.ncols;
        }
    }
    return 0;
}

}  // namespace tmb

extern "C" {

int64_t tm_sizeof_block_desc(void) { return (int64_t)sizeof(tm_block_desc); }

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
        off += tmb::self_elems(blocks[i]);
        for (int j = i + 1; j < n_blocks; ++j) off += blocks[i].ncols * blocks[j].ncols;
    }
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
