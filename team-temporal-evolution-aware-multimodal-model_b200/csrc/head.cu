// Host orchestration + C ABI of the fusion head: team_head_tri_fwd / team_head_tri_bwd /
// team_head_encode (include/team_b200.h).  Every launch goes to the caller's stream, there is
// no synchronisation and no allocation, so a whole step can be captured into a CUDA graph.
#include "head_bwd_kernels.cuh"
#include "gemm_tc.cuh"

namespace team {

HeadDims head_dims(int64_t B, int C, int P, int Tc) {
    HeadDims d;
    d.B = (int)B; d.B2 = 2 * (int)B; d.C = C; d.P = P; d.M = C + P; d.Ns = d.M + 10;
    d.Nsp = (d.Ns + 15) / 16 * 16; d.Rt = C + 10; d.Tc = Tc; d.nctas = NUM_SMS;
    return d;
}

static size_t tc_operand_bytes(const HeadDims& d) {
    // generous bound on the bf16 copies one fwd or bwd call makes (see BfCache below)
    const size_t B2 = d.B2, Nsp = d.Nsp;
    const size_t elems = B2 * (20 * (size_t)D + 8 * Nsp) + Nsp * (16 * (size_t)D + 6 * Nsp) + 16 * (size_t)D * D +
                         (size_t)(d.Rt + d.C + 64 + (d.Tc > 0 ? d.Tc : 0)) * 4 * D;
    return align_up(elems * 2 + 96 * 256, 256);
}

void head_plan(const HeadDims& d, int mode, void* base, HeadWS* w) {
    size_t off = 0;
    auto take = [&](size_t n_floats) -> float* {
        float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
        off += align_up(n_floats * sizeof(float), 256);
        return p;
    };
    const size_t B2 = d.B2, Nsp = d.Nsp, Rt = d.Rt;
    for (int i = 0; i < 3; ++i) { w->Wsum[i] = take((size_t)D * D); w->bsum[i] = take(D); }
    w->Ztab = take(Rt * D);
    w->S = take(Nsp * D); w->invS = take(Nsp);
    w->QKVs = take(Nsp * 3 * D); w->VFs = take(Nsp * D);
    w->TT = take(Nsp * Nsp); w->mt = take(Nsp); w->Zt = take(Nsp); w->Pt = take(Nsp * Nsp); w->NFt = take(Nsp * D);
    w->Xo = take(B2 * D); w->invo = take(B2);
    w->QKVo = take(B2 * 3 * D); w->VFo = take(B2 * D);
    w->SQ = take(B2 * Nsp); w->SK = take(B2 * Nsp);
    w->Aext = take(B2 * Nsp); w->aown = take(B2 * 2);
    w->Ybo = take(B2 * D); w->lnstat = take(B2 * 2);
    w->dYo = take(B2 * D); w->rowdot = take(B2); w->dsown = take(B2 * 2);
    w->dSK = take(B2 * Nsp); w->dVFo = take(B2 * D);
    w->dQKVo = take(B2 * 3 * D); w->dXo = take(B2 * D);
    w->Rfull = take(Nsp * D); w->Gfull = take(Nsp * D); w->hfull = take(Nsp);
    w->dTT = take(Nsp * Nsp); w->tmpNN = take(Nsp * Nsp); w->dVFs = take(Nsp * D);
    w->dQKVs = take(Nsp * 3 * D); w->dZtab = take(Rt * D);
    const TabOff to = tab_offsets(d);
    w->tab_partials = take((size_t)d.nctas * to.len); w->tab_reduced = take(to.len);
    w->own_partials = take((size_t)d.nctas * OWN_PARTIAL_LEN); w->own_reduced = take(OWN_PARTIAL_LEN);
    w->colsum_partials = take((size_t)64 * D);
    // split-K scratch: largest user is a [512,512] weight gradient reduced over 2B rows
    size_t g = gemm_f32_workspace_bytes(D, D, B2);
    const size_t g2 = gemm_f32_workspace_bytes(Nsp, D, B2);
    if (g2 > g) g = g2;
    if (g < (size_t)64 * D * D * sizeof(float)) g = (size_t)64 * D * D * sizeof(float);
    w->gemm_ws_bytes = g;
    w->gemm_ws = take(g / sizeof(float));
    w->bf16_bytes = (mode == TEAM_MODE_BF16) ? tc_operand_bytes(d) : 0;
    w->bf16_area = w->bf16_bytes ? (void*)take(w->bf16_bytes / sizeof(float)) : nullptr;
    // last, so that the layout of everything above does not depend on the number of text-class rows
    w->Zc = take((size_t)(d.Tc > 0 ? d.Tc : 1) * D);
    w->total_bytes = off;
}

// bf16 shadow copies of GEMM operands (mode BF16), valid for the duration of one fwd or bwd call.
// An operand is converted the first time a GEMM asks for it; sub-views of an already converted
// parent (same leading dimension) are served from the parent.  The orchestration below only asks
// for an operand after its final write, so no invalidation is needed.
struct BfCache {
    struct Ent { const float* base; int64_t rows, cols, ld; char* bf; };
    Ent e[96];
    int n = 0;
    char* area = nullptr;
    size_t cap = 0, used = 0;
};

struct HeadCtx {
    cudaStream_t st;
    int mode;
    HeadDims d;
    HeadWS w;
    BfCache bc;
};

static int bf16_view(HeadCtx& cx, const float* p, int64_t rows, int64_t cols, int64_t ld, const void** out) {
    BfCache& c = cx.bc;
    for (int i = 0; i < c.n; ++i) {
        const BfCache::Ent& e = c.e[i];
        if (e.ld != ld || p < e.base) continue;
        const int64_t off = p - e.base;
        const int64_t r0 = off / ld, c0 = off % ld;
        if (r0 + rows <= e.rows && c0 + cols <= e.cols) { *out = e.bf + off * 2; return TEAM_OK; }
    }
    const size_t bytes = align_up((size_t)((rows - 1) * ld + cols) * 2, 256);
    if (c.n >= 96 || c.used + bytes > c.cap) {
        set_error("head: bf16 operand staging exhausted (%zu + %zu > %zu, %d entries)", c.used, bytes, c.cap, c.n);
        return TEAM_EWORKSPACE;
    }
    char* dst = c.area + c.used;
    c.used += bytes;
    int rc = to_bf16(cx.st, p, ld, rows, (int)cols, dst, nullptr, ld);
    if (rc) return rc;
    c.e[c.n++] = BfCache::Ent{p, rows, cols, ld, dst};
    *out = dst;
    return TEAM_OK;
}

// convert a whole parent buffer now (its column sub-views are then served from it)
static int bf16_parent(HeadCtx& cx, const float* p, int64_t rows, int64_t cols) {
    if (cx.mode != TEAM_MODE_BF16) return TEAM_OK;
    const void* unused;
    return bf16_view(cx, p, rows, cols, cols, &unused);
}

// C[M,N] = alpha op(A) op(B) + beta C (+bias).  ta: A stored [K,M]; tb: B stored [N,K].
// F32 mode: fp32 FFMA GEMM.  BF16 mode: operands rounded to bf16 once, tcgen05 GEMM, fp32 accumulate.
static int hgemm(HeadCtx& cx, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias) {
    const bool tc_ok = cx.mode == TEAM_MODE_BF16 && (lda % 8 == 0) && (ldb % 8 == 0) && K >= 1 &&
                       ((reinterpret_cast<uintptr_t>(A) & 31) == 0) && ((reinterpret_cast<uintptr_t>(B) & 31) == 0) &&
                       ((ta ? M : K) % 4 == 0) && ((tb ? K : N) % 4 == 0);
    if (!tc_ok) return gemm_f32(cx.st, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, cx.w.gemm_ws, cx.w.gemm_ws_bytes);
    TcGemm g;
    g.a_mn = ta; g.b_mn = !tb; g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A2 = nullptr; g.Cb = nullptr; g.ldcb = 0; g.lda = lda; g.ldb = ldb; g.C = C; g.ldc = ldc; g.bias = bias;
    int rc = bf16_view(cx, A, ta ? K : M, ta ? M : K, lda, &g.A);
    if (rc) return rc;
    if ((rc = bf16_view(cx, B, tb ? N : K, tb ? K : N, ldb, &g.B))) return rc;
    return gemm_bf16_tc(cx.st, g, cx.w.gemm_ws, cx.w.gemm_ws_bytes);
}

static int validate(const team_head_weights* hw, int mode, int64_t batch) {
    TEAM_REQUIRE(hw != nullptr, "head: null weights");
    TEAM_REQUIRE(mode == TEAM_MODE_F32 || mode == TEAM_MODE_BF16, "head: bad mode %d", mode);
    TEAM_REQUIRE(hw->num_tasks >= 1 && hw->num_tasks <= TEAM_MAX_TASKS, "head: num_tasks %d out of [1,%d]", hw->num_tasks, TEAM_MAX_TASKS);
    TEAM_REQUIRE(hw->prompts_per_task >= 0 && hw->num_classes >= 1, "head: bad prompts_per_task/num_classes");
    TEAM_REQUIRE(batch >= 1 && batch < (1 << 22), "head: batch %lld out of range", (long long)batch);
    TEAM_REQUIRE(hw->num_classes + hw->num_tasks * hw->prompts_per_task <= 4096, "head: too many shared rows");
    return TEAM_OK;
}

static PtrList plist(const float* const* p, int n) {
    PtrList l;
    memset(&l, 0, sizeof(l));
    for (int i = 0; i < n; ++i) l.p[i] = p[i];
    l.n = n;
    return l;
}

#define HG(...)                                       \
    do {                                              \
        int _rc = hgemm(cx, __VA_ARGS__);             \
        if (_rc != TEAM_OK) return _rc;               \
    } while (0)

static int sum_projections(HeadCtx& cx, const team_head_weights* hw) {
    const int T = hw->num_tasks;
    const float* const* Ws[3] = {hw->w_img, hw->w_text, hw->w_state};
    const float* const* Bs[3] = {hw->b_img, hw->b_text, hw->b_state};
    for (int k = 0; k < 3; ++k) {
        sum_weights_kernel<<<D * D / 4 / 256, 256, 0, cx.st>>>(plist(Ws[k], T), plist(Bs[k], T), cx.w.Wsum[k], cx.w.bsum[k]);
        TEAM_LAUNCH_CHECK("sum_weights_kernel");
    }
    return TEAM_OK;
}

// X[M,D] @ {Wq,Wk,Wv}^T -> out[M,3D]
static int qkv_forward(HeadCtx& cx, const team_head_weights* hw, const float* X, int64_t rows, float* out) {
    HG(false, true, rows, D, D, 1.f, X, D, hw->w_q, D, 0.f, out, 3 * D, nullptr);
    HG(false, true, rows, D, D, 1.f, X, D, hw->w_k, D, 0.f, out + D, 3 * D, nullptr);
    HG(false, true, rows, D, D, 1.f, X, D, hw->w_v, D, 0.f, out + 2 * D, 3 * D, nullptr);
    return TEAM_OK;
}

static int step_rows_forward(HeadCtx& cx, const team_head_weights* hw) {
    const HeadDims& d = cx.d;
    HeadWS& w = cx.w;
    // prototype rows and state-table rows: project, normalise, place into S
    HG(false, true, d.C, D, D, 1.f, hw->prototypes, D, w.Wsum[0], D, 0.f, w.Ztab, D, w.bsum[0]);
    HG(false, true, 10, D, D, 1.f, hw->state_emb, D, w.Wsum[2], D, 0.f, w.Ztab + (size_t)d.C * D, D, w.bsum[2]);
    rows_normalize_kernel<<<(d.C + 7) / 8, 256, 0, cx.st>>>(w.Ztab, d.C, w.S, w.invS, 1);
    TEAM_LAUNCH_CHECK("rows_normalize_kernel");
    rows_normalize_kernel<<<2, 256, 0, cx.st>>>(w.Ztab + (size_t)d.C * D, 10, w.S + (size_t)d.M * D, w.invS + d.M, 1);
    TEAM_LAUNCH_CHECK("rows_normalize_kernel");
    const int fill_rows = d.P + (d.Nsp - d.Ns);
    if (fill_rows > 0) {
        fill_prompt_rows_kernel<<<fill_rows, 128, 0, cx.st>>>(plist(hw->prompts, hw->num_tasks), hw->prompts_per_task > 0 ? hw->prompts_per_task : 1, d.C, d.Ns, d.Nsp, w.S);
        TEAM_LAUNCH_CHECK("fill_prompt_rows_kernel");
    }
    int rc = qkv_forward(cx, hw, w.S, d.Nsp, w.QKVs);
    if (rc) return rc;
    if ((rc = bf16_parent(cx, w.QKVs, d.Nsp, 3 * D))) return rc;
    HG(false, true, d.Nsp, D, D, 1.f, w.QKVs + 2 * D, 3 * D, hw->w_fc, D, 0.f, w.VFs, D, nullptr);
    HG(false, true, d.Nsp, d.Nsp, D, 1.f, w.QKVs, 3 * D, w.QKVs + D, 3 * D, 0.f, w.TT, d.Nsp, nullptr);
    table_prep_kernel<<<(d.Nsp + 7) / 8, 256, 0, cx.st>>>(w.TT, d.M, d.Nsp, w.mt, w.Zt, w.Pt);
    TEAM_LAUNCH_CHECK("table_prep_kernel");
    HG(false, false, d.Nsp, D, d.Nsp, 1.f, w.Pt, d.Nsp, w.VFs, D, 0.f, w.NFt, D, nullptr);
    return TEAM_OK;
}

static int setup(HeadCtx& cx, const team_head_weights* hw, int mode, int64_t batch, int Tc, void* workspace,
                 size_t workspace_bytes, void* stream) {
    int rc = validate(hw, mode, batch);
    if (rc) return rc;
    cx.st = (cudaStream_t)stream;
    cx.mode = mode;
    cx.d = head_dims(batch, hw->num_classes, hw->num_tasks * hw->prompts_per_task, Tc);
    head_plan(cx.d, mode, workspace, &cx.w);
    if (workspace == nullptr || workspace_bytes < cx.w.total_bytes) {
        set_error("head: workspace %zu < %zu bytes", workspace_bytes, cx.w.total_bytes);
        return TEAM_EWORKSPACE;
    }
    TEAM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "head: workspace must be 256-byte aligned");
    cx.bc.n = 0; cx.bc.used = 0;
    cx.bc.area = reinterpret_cast<char*>(cx.w.bf16_area);
    cx.bc.cap = cx.w.bf16_bytes;
    if (mode == TEAM_MODE_BF16) return tc_workspace_init(cx.st, cx.w.gemm_ws, cx.w.gemm_ws_bytes);
    return TEAM_OK;
}

}  // namespace team

using namespace team;

extern "C" size_t team_head_workspace_bytes(int64_t batch, int32_t num_classes, int32_t num_prompts,
                                            int32_t num_text_cls, int mode) {
    HeadDims d = head_dims(batch, num_classes, num_prompts, num_text_cls);
    HeadWS w;
    head_plan(d, mode, nullptr, &w);
    return w.total_bytes;
}

extern "C" int team_head_tri_fwd(const team_head_weights* hw, int mode, int64_t batch, const float* image_feat,
                                 const float* text_feat, const int64_t* state_ids, const float* text_cls,
                                 int64_t num_text_cls, float* out_image, float* out_text, float* out_state,
                                 float* out_proto, float* cls_logits, int64_t* cls_argmax, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    HeadCtx cx;
    int rc = setup(cx, hw, mode, batch, (int)(text_cls ? num_text_cls : 0), workspace, workspace_bytes, stream);
    if (rc) return rc;
    TEAM_REQUIRE(image_feat && text_feat && state_ids && out_image && out_text && out_state && out_proto, "head fwd: null pointer");
    const HeadDims& d = cx.d;
    HeadWS& w = cx.w;
    if ((rc = sum_projections(cx, hw))) return rc;
    if ((rc = step_rows_forward(cx, hw))) return rc;
    // own rows: image rows [0,B), text rows [B,2B)
    HG(false, true, d.B, D, D, 1.f, image_feat, D, w.Wsum[0], D, 0.f, w.Xo, D, w.bsum[0]);
    HG(false, true, d.B, D, D, 1.f, text_feat, D, w.Wsum[1], D, 0.f, w.Xo + (size_t)d.B * D, D, w.bsum[1]);
    rows_normalize_kernel<<<(d.B2 + 7) / 8, 256, 0, cx.st>>>(w.Xo, d.B2, w.Xo, w.invo, 1);
    TEAM_LAUNCH_CHECK("rows_normalize_kernel");
    if ((rc = qkv_forward(cx, hw, w.Xo, d.B2, w.QKVo))) return rc;
    if ((rc = bf16_parent(cx, w.QKVo, d.B2, 3 * D))) return rc;
    HG(false, true, d.B2, D, D, 1.f, w.QKVo + 2 * D, 3 * D, hw->w_fc, D, 0.f, w.VFo, D, nullptr);
    HG(false, true, d.B2, d.Nsp, D, 1.f, w.QKVo, 3 * D, w.QKVs + D, 3 * D, 0.f, w.SQ, d.Nsp, nullptr);
    HG(false, true, d.B2, d.Nsp, D, 1.f, w.QKVo + D, 3 * D, w.QKVs, 3 * D, 0.f, w.SK, d.Nsp, nullptr);
    attn_own_kernel<<<(d.B2 + 7) / 8, 256, 0, cx.st>>>(d, w.SQ, w.QKVo, state_ids, w.Aext, w.aown);
    TEAM_LAUNCH_CHECK("attn_own_kernel");
    HG(false, false, d.B2, D, d.Nsp, 1.f, w.Aext, d.Nsp, w.VFs, D, 0.f, w.Ybo, D, nullptr);
    ln_own_fwd_kernel<<<(d.B2 + 7) / 8, 256, 0, cx.st>>>(d, w.Ybo, w.aown, w.VFo, w.Xo, hw->b_fc, hw->ln_g, hw->ln_b, w.lnstat, out_image, out_text);
    TEAM_LAUNCH_CHECK("ln_own_fwd_kernel");
    const int tgrid = d.B < 2 * NUM_SMS ? d.B : 2 * NUM_SMS;
    table_rows_fwd_kernel<<<tgrid, TR_WARPS * 32, 0, cx.st>>>(d, w.SK, w.TT, w.mt, w.Zt, w.NFt, w.VFo, w.VFs, w.S, hw->b_fc, hw->ln_g, hw->ln_b, state_ids, out_proto, out_state);
    TEAM_LAUNCH_CHECK("table_rows_fwd_kernel");
    if (text_cls != nullptr && num_text_cls > 0 && (cls_logits != nullptr || cls_argmax != nullptr)) {
        // forward_for_classification (models/proof.py:519-536): image rows are already normalised in Xo
        HG(false, true, num_text_cls, D, D, 1.f, text_cls, D, w.Wsum[1], D, 0.f, w.Zc, D, w.bsum[1]);
        if ((rc = cosine_logits_launch(cx.st, w.Xo, d.B, w.Zc, num_text_cls, nullptr, cls_logits, cls_argmax))) return rc;
    }
    return TEAM_OK;
}

static int colsum(HeadCtx& cx, const float* X, int64_t rows, float* out, int accumulate) {
    int chunks = (int)((rows + 31) / 32);
    if (chunks > 64) chunks = 64;
    if (chunks < 1) chunks = 1;
    const int64_t rpc = (rows + chunks - 1) / chunks;
    colsum_partial_kernel<<<chunks, 128, 0, cx.st>>>(X, rows, rpc, cx.w.colsum_partials);
    TEAM_LAUNCH_CHECK("colsum_partial_kernel");
    colsum_final_kernel<<<1, 128, 0, cx.st>>>(cx.w.colsum_partials, chunks, out, accumulate);
    TEAM_LAUNCH_CHECK("colsum");
    return TEAM_OK;
}

extern "C" int team_head_tri_bwd(const team_head_weights* hw, int mode, int64_t batch, const float* image_feat,
                                 const float* text_feat, const int64_t* state_ids, const float* g_image,
                                 const float* g_text, const float* g_state, const float* g_proto,
                                 const team_head_grads* gr, void* workspace, size_t workspace_bytes, void* stream) {
    HeadCtx cx;
    int rc = setup(cx, hw, mode, batch, 0, workspace, workspace_bytes, stream);
    if (rc) return rc;
    TEAM_REQUIRE(gr && g_image && g_text && g_state && g_proto && image_feat && text_feat && state_ids, "head bwd: null pointer");
    TEAM_REQUIRE(gr->w_img && gr->b_img && gr->w_text && gr->b_text && gr->w_state && gr->b_state && gr->state_emb &&
                 gr->w_q && gr->w_k && gr->w_v && gr->w_fc && gr->b_fc && gr->ln_g && gr->ln_b, "head bwd: null gradient buffer");
    const HeadDims& d = cx.d;
    HeadWS& w = cx.w;
    const TabOff to = tab_offsets(d);
    if ((rc = bf16_parent(cx, w.QKVo, d.B2, 3 * D))) return rc;
    if ((rc = bf16_parent(cx, w.QKVs, d.Nsp, 3 * D))) return rc;
    // ---- table-query rows (prototype / state outputs)
    const int tgrid = d.B < d.nctas ? d.B : d.nctas;
    const size_t tsm = (size_t)(3 * TR_WARPS * D + 10 * D + 3 * D + d.Rt * 11) * sizeof(float);
    TEAM_CUDA_CHECK(cudaFuncSetAttribute(table_rows_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
    table_rows_bwd_kernel<<<tgrid, TR_WARPS * 32, tsm, cx.st>>>(d, w.SK, w.TT, w.mt, w.Zt, w.NFt, w.VFo, w.VFs, w.S, hw->b_fc, hw->ln_g, hw->ln_b, state_ids, g_proto, g_state, w.dSK, w.dVFo, w.tab_partials);
    TEAM_LAUNCH_CHECK("table_rows_bwd_kernel");
    reduce_partials_kernel<<<(unsigned)((to.len / 4 + 63) / 64), 256, 0, cx.st>>>(w.tab_partials, tgrid, to.len / 4, w.tab_reduced);
    TEAM_LAUNCH_CHECK("reduce_partials_kernel");
    expand_table_kernel<<<d.Nsp, 128, 0, cx.st>>>(d, w.tab_reduced, w.Rfull, w.Gfull, w.hfull, w.dTT, w.dVFs);
    TEAM_LAUNCH_CHECK("expand_table_kernel");
    // ---- own query rows
    int ogrid = (d.B + 7) / 8;
    if (ogrid > d.nctas) ogrid = d.nctas;
    ln_own_bwd_kernel<<<ogrid, 256, 0, cx.st>>>(d, w.Ybo, w.Xo, w.VFo, w.aown, hw->b_fc, hw->ln_g, hw->ln_b, g_image, g_text, w.dYo, w.dXo, w.rowdot, w.dsown, w.dVFo, w.own_partials);
    TEAM_LAUNCH_CHECK("ln_own_bwd_kernel");
    reduce_partials_kernel<<<(OWN_PARTIAL_LEN / 4 + 63) / 64, 256, 0, cx.st>>>(w.own_partials, ogrid, OWN_PARTIAL_LEN / 4, w.own_reduced);
    TEAM_LAUNCH_CHECK("reduce_partials_kernel");
    finalize_ln_grads_kernel<<<1, 128, 0, cx.st>>>(d, w.tab_reduced, w.own_reduced, gr->ln_g, gr->ln_b, gr->b_fc);
    TEAM_LAUNCH_CHECK("ln_own_bwd");
    // dA = dYo VFs^T  (into the SQ buffer), dVFs += Aext^T dYo, dS = Aext.*(dA - rowdot)/tau
    HG(false, true, d.B2, d.Nsp, D, 1.f, w.dYo, D, w.VFs, D, 0.f, w.SQ, d.Nsp, nullptr);
    HG(true, false, d.Nsp, D, d.B2, 1.f, w.Aext, d.Nsp, w.dYo, D, 1.f, w.dVFs, D, nullptr);
    {
        const int64_t n = (int64_t)d.B2 * d.Nsp;
        ds_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx.st>>>(n, d.Nsp, w.Aext, w.rowdot, w.SQ);
        TEAM_LAUNCH_CHECK("ds_kernel");
    }
    float* dS = w.SQ;
    float *dQo = w.dQKVo, *dKo = w.dQKVo + D, *dVo = w.dQKVo + 2 * D;
    float *dQs = w.dQKVs, *dKs = w.dQKVs + D, *dVs = w.dQKVs + 2 * D;
    const float *Qo = w.QKVo, *Ko = w.QKVo + D, *Vo = w.QKVo + 2 * D;
    const float *Qs = w.QKVs, *Ks = w.QKVs + D, *Vs = w.QKVs + 2 * D;
    HG(false, false, d.B2, D, d.Nsp, 1.f, dS, d.Nsp, Ks, 3 * D, 0.f, dQo, 3 * D, nullptr);          // dQo = dS Ks
    HG(true, false, d.Nsp, D, d.B2, 1.f, dS, d.Nsp, Qo, 3 * D, 0.f, dKs, 3 * D, nullptr);           // dKs = dS^T Qo
    HG(false, false, d.B2, D, d.Nsp, 1.f, w.dSK, d.Nsp, Qs, 3 * D, 0.f, dKo, 3 * D, nullptr);       // dKo = dSK Qs
    HG(true, false, d.Nsp, D, d.B2, 1.f, w.dSK, d.Nsp, Ko, 3 * D, 0.f, dQs, 3 * D, nullptr);        // dQs = dSK^T Ko
    own_own_bwd_kernel<<<(d.B + 7) / 8, 256, 0, cx.st>>>(d, w.QKVo, w.dsown, w.dQKVo);
    TEAM_LAUNCH_CHECK("own_own_bwd_kernel");
    // ---- per-step part of the table queries
    HG(false, true, d.Nsp, d.Nsp, D, 1.f, w.Gfull, D, w.VFs, D, 0.f, w.tmpNN, d.Nsp, nullptr);     // G VFs^T
    dtt_kernel<<<(d.Nsp * d.Nsp + 255) / 256, 256, 0, cx.st>>>(d.Nsp, d.M, w.Pt, w.tmpNN, w.hfull, w.dTT);
    TEAM_LAUNCH_CHECK("dtt_kernel");
    HG(true, false, d.Nsp, D, d.Nsp, 1.f, w.Pt, d.Nsp, w.Gfull, D, 1.f, w.dVFs, D, nullptr);        // dVFs += P^T G
    HG(false, false, d.Nsp, D, d.Nsp, 1.f, w.dTT, d.Nsp, Ks, 3 * D, 1.f, dQs, 3 * D, nullptr);      // dQs += dTT Ks
    HG(true, false, d.Nsp, D, d.Nsp, 1.f, w.dTT, d.Nsp, Qs, 3 * D, 1.f, dKs, 3 * D, nullptr);       // dKs += dTT^T Qs
    // ---- fc folded into V
    HG(false, false, d.B2, D, D, 1.f, w.dVFo, D, hw->w_fc, D, 0.f, dVo, 3 * D, nullptr);            // dVo = dVFo Wfc
    HG(false, false, d.Nsp, D, D, 1.f, w.dVFs, D, hw->w_fc, D, 0.f, dVs, 3 * D, nullptr);
    HG(true, false, D, D, d.B2, 1.f, w.dVFo, D, Vo, 3 * D, 0.f, gr->w_fc, D, nullptr);              // dWfc = dVFo^T Vo
    HG(true, false, D, D, d.Nsp, 1.f, w.dVFs, D, Vs, 3 * D, 1.f, gr->w_fc, D, nullptr);             //      + dVFs^T Vs
    // ---- q/k/v projections
    if ((rc = bf16_parent(cx, w.dQKVo, d.B2, 3 * D))) return rc;
    if ((rc = bf16_parent(cx, w.dQKVs, d.Nsp, 3 * D))) return rc;
    const float* Wqkv[3] = {hw->w_q, hw->w_k, hw->w_v};
    float* dWqkv[3] = {gr->w_q, gr->w_k, gr->w_v};
    for (int i = 0; i < 3; ++i) {
        HG(false, false, d.B2, D, D, 1.f, w.dQKVo + i * D, 3 * D, Wqkv[i], D, 1.f, w.dXo, D, nullptr);      // dXo += dQ Wq ...
        HG(false, false, d.Nsp, D, D, 1.f, w.dQKVs + i * D, 3 * D, Wqkv[i], D, 1.f, w.Rfull, D, nullptr);   // dS_rows (in Rfull)
        HG(true, false, D, D, d.B2, 1.f, w.dQKVo + i * D, 3 * D, w.Xo, D, 0.f, dWqkv[i], D, nullptr);
        HG(true, false, D, D, d.Nsp, 1.f, w.dQKVs + i * D, 3 * D, w.S, D, 1.f, dWqkv[i], D, nullptr);
    }
    // ---- normalisations and the newest projections
    nrm_bwd_kernel<<<(d.B2 + 7) / 8, 256, 0, cx.st>>>(w.dXo, d.B2, w.Xo, w.invo, nullptr, d.B2, 0);
    TEAM_LAUNCH_CHECK("nrm_bwd_kernel");
    // proto rows [0,C) and state rows [M,M+10) of dS_rows -> compact dZtab [C+10]
    nrm_bwd_kernel<<<(d.Rt + 7) / 8, 256, 0, cx.st>>>(w.dZtab, d.Rt, w.S, w.invS, w.Rfull, d.C, d.P);
    TEAM_LAUNCH_CHECK("nrm_bwd_kernel");
    const float* dz0 = w.dXo;
    const float* dz1 = w.dXo + (size_t)d.B * D;
    const float* dzp = w.dZtab;
    const float* dzs = w.dZtab + (size_t)d.C * D;
    HG(true, false, D, D, d.B, 1.f, dz0, D, image_feat, D, 0.f, gr->w_img, D, nullptr);              // dWi = dz0^T x
    HG(true, false, D, D, d.C, 1.f, dzp, D, hw->prototypes, D, 1.f, gr->w_img, D, nullptr);          //     + dzp^T protos
    HG(true, false, D, D, d.B, 1.f, dz1, D, text_feat, D, 0.f, gr->w_text, D, nullptr);
    HG(true, false, D, D, 10, 1.f, dzs, D, hw->state_emb, D, 0.f, gr->w_state, D, nullptr);
    HG(false, false, 10, D, D, 1.f, dzs, D, w.Wsum[2], D, 0.f, gr->state_emb, D, nullptr);           // dE = dzs Ws
    if ((rc = colsum(cx, dz0, d.B, gr->b_img, 0))) return rc;
    if ((rc = colsum(cx, dzp, d.C, gr->b_img, 1))) return rc;
    if ((rc = colsum(cx, dz1, d.B, gr->b_text, 0))) return rc;
    if ((rc = colsum(cx, dzs, 10, gr->b_state, 0))) return rc;
    if (gr->prompts != nullptr && d.P > 0) {   // gradient of every prompt row (callers keep the newest task's slice)
        TEAM_CUDA_CHECK(cudaMemcpyAsync(gr->prompts, w.Rfull + (size_t)d.C * D, (size_t)d.P * D * sizeof(float), cudaMemcpyDeviceToDevice, cx.st));
    }
    return TEAM_OK;
}

extern "C" int team_head_encode(const team_head_weights* hw, int mode, int which, const void* x, int64_t n_rows,
                                int normalize, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    HeadCtx cx;
    int rc = setup(cx, hw, mode, 1, 0, workspace, workspace_bytes, stream);
    if (rc) return rc;
    TEAM_REQUIRE(which >= 0 && which <= 3 && out != nullptr, "head encode: bad args");
    if ((rc = sum_projections(cx, hw))) return rc;
    const int k = which == 3 ? 0 : which;
    const float* src = reinterpret_cast<const float*>(x);
    if (which == 3) { src = hw->prototypes; n_rows = hw->num_classes; }
    if (which == 2) {
        // 10-row table first, then gather by state id
        TEAM_REQUIRE(x != nullptr, "head encode: null state ids");
        HG(false, true, 10, D, D, 1.f, hw->state_emb, D, cx.w.Wsum[2], D, 0.f, cx.w.Ztab, D, cx.w.bsum[2]);
        rows_normalize_kernel<<<2, 256, 0, cx.st>>>(cx.w.Ztab, 10, cx.w.Ztab, nullptr, normalize);
        TEAM_LAUNCH_CHECK("rows_normalize_kernel");
        gather_rows_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, cx.st>>>(cx.w.Ztab, reinterpret_cast<const int64_t*>(x), n_rows, out);
        TEAM_LAUNCH_CHECK("gather_rows_kernel");
        return TEAM_OK;
    }
    TEAM_REQUIRE(src != nullptr, "head encode: null input");
    HG(false, true, n_rows, D, D, 1.f, src, D, cx.w.Wsum[k], D, 0.f, out, D, cx.w.bsum[k]);
    if (normalize) {
        rows_normalize_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, cx.st>>>(out, n_rows, out, nullptr, 1);
        TEAM_LAUNCH_CHECK("rows_normalize_kernel");
    }
    return TEAM_OK;
}
