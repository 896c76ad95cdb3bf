// Contrastive losses of the training step, forward + gradient in one call (SURVEY 8f "next" rows):
//   team_unicl_loss : unicl_loss (models/proof.py:21-191, evolution_features = None) - the three-way contrastive
//                     loss on the head outputs; its gradients ARE the cotangents of team_head_tri_bwd
//   team_clip_loss  : ClipLoss.forward (utils/toolkit.py:128-141, world_size 1) on the encode_image / encode_text rows
// The B x B similarity matrices and the two "coefficient x rows" gradient products are grouped tensor-core GEMMs
// (fp32 FFMA GEMMs in TEAM_MODE_F32); the softmax statistics, the per-sample 3 x 3 instance term and the
// normalise-backward are warp-per-row kernels; every batch reduction is a fixed-order fold (deterministic).
#include "head_kernels.cuh"
#include "gemm_tc.cuh"

namespace team {

constexpr int64_t LOSS_MAX_BATCH = 16384;        // the losses are O(B^2) by definition (B x B similarities)

struct LossWS {
    Mat X[3];                 // normalised image / text / state rows [B][D]
    float* inv[3];            // 1 / max(|x|, eps)
    float* sim;               // [B][B]
    float* simT;              // [B][B] (clip only)
    Mat G;                    // [B][B] gradient coefficients
    float* rmax; float* rpos; float* rall; float* rloss;     // [2B]
    float* dcat;              // [B][D]
    float* part;              // [1024] fold scratch
    float* scal;              // [8] device scalars: 0 valid count, 1 category loss, 2 instance sum
    void* gemm_ws; size_t gemm_ws_bytes;
    size_t total;
};

static void loss_plan(int64_t B, void* base, LossWS* w) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> char* {
        char* p = base ? reinterpret_cast<char*>(base) + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    for (int i = 0; i < 3; ++i) {
        w->X[i].f = reinterpret_cast<float*>(take((size_t)B * D * 4));
        w->X[i].h = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * D * 2));
        w->X[i].ld = D;
        w->inv[i] = reinterpret_cast<float*>(take((size_t)B * 4));
    }
    const int64_t Bp = (B + 7) / 8 * 8;                      // leading dimension of the B x B matrices
    w->sim = reinterpret_cast<float*>(take((size_t)B * Bp * 4));
    w->simT = reinterpret_cast<float*>(take((size_t)B * Bp * 4));
    w->G.f = reinterpret_cast<float*>(take((size_t)B * Bp * 4));
    w->G.h = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * Bp * 2));
    w->G.ld = Bp;
    w->rmax = reinterpret_cast<float*>(take((size_t)2 * B * 4));
    w->rpos = reinterpret_cast<float*>(take((size_t)2 * B * 4));
    w->rall = reinterpret_cast<float*>(take((size_t)2 * B * 4));
    w->rloss = reinterpret_cast<float*>(take((size_t)2 * B * 4));
    w->dcat = reinterpret_cast<float*>(take((size_t)B * D * 4));
    w->part = reinterpret_cast<float*>(take(1024 * 4));
    w->scal = reinterpret_cast<float*>(take(8 * 4));
    w->gemm_ws_bytes = 64u << 20;
    w->gemm_ws = take(w->gemm_ws_bytes);
    w->total = off;
}

// one product C[M,N] = op(A) op(B) (+ second K-segment) through the mode's GEMM engine
struct LSeg { bool a_mn, b_mn; int64_t K; Mat A, B; };
static int loss_gemm(cudaStream_t st, int mode, LossWS& w, int64_t M, int64_t N, float* C, int64_t ldc, const LSeg* s, int nseg) {
    if (mode == TEAM_MODE_BF16) {
        TcGemm g;
        memset(&g, 0, sizeof(g));
        g.M = M; g.N = N; g.nseg = nseg; g.alpha = 1.f; g.beta = 0.f; g.C = C; g.ldc = ldc;
        for (int q = 0; q < nseg; ++q) {
            g.s[q].a_mn = s[q].a_mn; g.s[q].b_mn = s[q].b_mn; g.s[q].K = s[q].K;
            g.s[q].A = s[q].A.h; g.s[q].lda = s[q].A.ld; g.s[q].B = s[q].B.h; g.s[q].ldb = s[q].B.ld;
        }
        return gemm_bf16_group(st, &g, 1, w.gemm_ws, w.gemm_ws_bytes);
    }
    for (int q = 0; q < nseg; ++q) {
        const int rc = gemm_f32(st, s[q].a_mn, !s[q].b_mn, M, N, s[q].K, 1.f, s[q].A.f, s[q].A.ld, s[q].B.f, s[q].B.ld,
                                q == 0 ? 0.f : 1.f, C, ldc, nullptr, w.gemm_ws, w.gemm_ws_bytes);
        if (rc) return rc;
    }
    return TEAM_OK;
}

// ------------------------------------------------------------------ kernels
// rows [0,B): image, [B,2B): text, [2B,3B): state -> normalised copies (+ bf16 shadow) and inverse norms
__global__ void __launch_bounds__(256)
loss_normalize_kernel(int64_t B, const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ x2,
                      float* y0, float* y1, float* y2, __nv_bfloat16* h0, __nv_bfloat16* h1, __nv_bfloat16* h2,
                      float* i0, float* i1, float* i2, int n_sets, int do_normalize) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= (int64_t)n_sets * B) return;
    const int k = (int)(r / B);
    const int64_t b = r - (int64_t)k * B;
    const float* x = (k == 0 ? x0 : k == 1 ? x1 : x2) + b * D;
    float* y = (k == 0 ? y0 : k == 1 ? y1 : y2) + b * D;
    __nv_bfloat16* h = k == 0 ? h0 : k == 1 ? h1 : h2;
    float* iv = k == 0 ? i0 : k == 1 ? i1 : i2;
    float4 v[4];
    ld_row(x, lane, v);
    float s = 1.f;
    if (do_normalize) {
        s = 1.0f / fmaxf(sqrtf(warp_sum(dot_part(v, v))), NORM_EPS);
        scale_row(v, s);
    }
    st_row(y, lane, v);
    st_row_h(h != nullptr ? h + b * D : nullptr, lane, v);
    if (lane == 0 && iv != nullptr) iv[b] = s;
}

// category term, per row i of sim / tau: max over ALL j, pos = sum_{j != i, y_j == y_i} e_j, all = sum_{j != i} e_j
__global__ void __launch_bounds__(256)
unicl_cat_stats_kernel(int B, int ld, const float* __restrict__ sim, const int64_t* __restrict__ labels, float inv_tau,
                       float* __restrict__ rmax, float* __restrict__ rpos, float* __restrict__ rall, float* __restrict__ rloss) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B) return;
    const float* row = sim + (size_t)i * ld;
    const int64_t yi = labels[i];
    float mx = -INFINITY;
    for (int j = lane; j < B; j += 32) mx = fmaxf(mx, row[j] * inv_tau);
    mx = warp_max(mx);
    float pos = 0.f, all = 0.f;
    for (int j = lane; j < B; j += 32) {
        if (j == i) continue;
        const float e = expf(row[j] * inv_tau - mx);
        all += e;
        if (labels[j] == yi) pos += e;
    }
    pos = warp_sum(pos); all = warp_sum(all);
    if (lane == 0) {
        const bool valid = pos > 0.f && all > 0.f;
        rmax[i] = mx; rpos[i] = pos; rall[i] = all;
        rloss[i] = valid ? -logf(pos / (all + 1e-8f)) : 0.f;
        rmax[B + i] = valid ? 1.f : 0.f;                      // valid flag
    }
}

// single block: fixed-order sums of n values of a and (optionally) b -> out[0], out[1]
__global__ void __launch_bounds__(1024)
loss_fold2_kernel(int n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sa[1024], sb[1024];
    float x = 0.f, y = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) { x += a[i]; if (b != nullptr) y += b[i]; }
    sa[threadIdx.x] = x; sb[threadIdx.x] = y;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sa[0]; out[1] = sb[0]; }
}

// G[i][j] = d(0.5 * grad_scale * category_loss) / d(sim_ij)   (scal[0] = valid count)
__global__ void __launch_bounds__(256)
unicl_cat_grad_kernel(int B, int ld, const float* __restrict__ sim, const int64_t* __restrict__ labels, float inv_tau,
                      float weight, const float* __restrict__ rmax, const float* __restrict__ rpos,
                      const float* __restrict__ rall, const float* __restrict__ scal, float* __restrict__ G,
                      __nv_bfloat16* __restrict__ Gh) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B) return;
    const float nvalid = scal[0];
    const bool valid = rmax[B + i] != 0.f && nvalid > 0.f;
    const float k = valid ? weight * inv_tau / nvalid : 0.f;
    const float mx = rmax[i], ip = valid ? 1.0f / rpos[i] : 0.f, ia = valid ? 1.0f / (rall[i] + 1e-8f) : 0.f;
    const int64_t yi = labels[i];
    const float* row = sim + (size_t)i * ld;
    for (int j = lane; j < ld; j += 32) {
        float g = 0.f;
        if (j < B && j != i && valid) {
            const float e = expf(row[j] * inv_tau - mx);
            g = k * e * (ia - (labels[j] == yi ? ip : 0.f));
        }
        G[(size_t)i * ld + j] = g;
        if (Gh != nullptr) Gh[(size_t)i * ld + j] = __float2bfloat16_rn(g);
    }
}

// instance term + assembly of the three cotangents (warp per sample):
//   3 x 3 similarities of (image, text, state) / tau; row r: pos = 1 + sum_{c != r} e_rc, all = sum_c e_rc,
//   loss -= log(pos / (all + 1e-8));  then d total / d xhat = instance part + dcat (image only), normalise-backward.
__global__ void __launch_bounds__(256)
unicl_instance_kernel(int B, const float* __restrict__ Xi, const float* __restrict__ Xt, const float* __restrict__ Xs,
                      const float* __restrict__ inv_i, const float* __restrict__ inv_t, const float* __restrict__ inv_s,
                      const float* __restrict__ dcat, float inv_tau, float inst_weight, float* __restrict__ g_image,
                      float* __restrict__ g_text, float* __restrict__ g_state, float* __restrict__ rloss) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    float4 v[3][4];
    ld_row(Xi + (size_t)b * D, lane, v[0]);
    ld_row(Xt + (size_t)b * D, lane, v[1]);
    ld_row(Xs + (size_t)b * D, lane, v[2]);
    float s[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = r; c < 3; ++c) {
            const float d = warp_sum(dot_part(v[r], v[c])) * inv_tau;
            s[r][c] = d; s[c][r] = d;
        }
    float loss = 0.f, wgt[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float e[3], pos = 1.f, all = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { e[c] = expf(s[r][c]); all += e[c]; if (c != r) pos += e[c]; }
        loss -= logf(pos / (all + 1e-8f));
        const float ia = 1.0f / (all + 1e-8f), ip = 1.0f / pos;
#pragma unroll
        for (int c = 0; c < 3; ++c) wgt[r][c] = c == r ? e[c] * ia : e[c] * (ia - ip);
    }
    if (lane == 0) rloss[b] = loss;
    const float k = inst_weight * inv_tau;
    const float* invs[3] = {inv_i, inv_t, inv_s};
    float* outs[3] = {g_image, g_text, g_state};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        // d/d v_r = sum_{c != r} (w_rc + w_cr) v_c + 2 w_rr v_r
        float4 dv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dv[i] = mul4s(2.f * k * wgt[r][r], v[r][i]);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (c != r) axpy_row(dv, k * (wgt[r][c] + wgt[c][r]), v[c]);
        if (r == 0) {
            float4 t[4];
            ld_row(dcat + (size_t)b * D, lane, t);
            add_row(dv, t);
        }
        const float proj = warp_sum(dot_part(dv, v[r]));
        const float iv = invs[r][b];
#pragma unroll
        for (int i = 0; i < 4; ++i) dv[i] = mul4s(iv, fma4s(-proj, v[r][i], dv[i]));
        st_row(outs[r] + (size_t)b * D, lane, dv);
    }
}

// losses[0] = total, [1] = instance, [2] = category   (scal: 0 valid, 1 category sum, 2 instance sum)
__global__ void unicl_finish_kernel(const float* __restrict__ scal, float inv_3B, float* __restrict__ losses) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) {
        const float cat = scal[0] > 0.f ? scal[1] / scal[0] : 0.f;
        const float inst = scal[2] * inv_3B;
        losses[0] = inst + 0.5f * cat; losses[1] = inst; losses[2] = cat;
    }
}

// ---- ClipLoss: per row of s * L (rows [0,B): image->text from L, rows [B,2B): text->image from L^T): lse and loss
__global__ void __launch_bounds__(256)
clip_stats_kernel(int B, int ld, const float* __restrict__ L, const float* __restrict__ LT, float scale,
                  float* __restrict__ lse, float* __restrict__ rloss) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= 2 * B) return;
    const int i = r < B ? r : r - B;
    const float* row = (r < B ? L : LT) + (size_t)i * ld;
    float mx = -INFINITY;
    for (int j = lane; j < B; j += 32) mx = fmaxf(mx, row[j] * scale);
    mx = warp_max(mx);
    float z = 0.f;
    for (int j = lane; j < B; j += 32) z += expf(row[j] * scale - mx);
    z = warp_sum(z);
    if (lane == 0) {
        const float l = logf(z) + mx;
        lse[r] = l;
        rloss[r] = l - row[i] * scale;
    }
}
// G[i][j] = d(grad_scale * loss) / d(L_ij) = k * (softmax_i2t[i][j] + softmax_t2i[j][i] - 2 delta_ij),  k = scale*grad_scale/(2B)
__global__ void __launch_bounds__(256)
clip_grad_kernel(int B, int ld, const float* __restrict__ L, float scale, float k, const float* __restrict__ lse,
                 float* __restrict__ G, __nv_bfloat16* __restrict__ Gh) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= B) return;
    const float li = lse[i];
    for (int j = lane; j < ld; j += 32) {
        float g = 0.f;
        if (j < B) {
            const float x = L[(size_t)i * ld + j] * scale;
            g = k * (expf(x - li) + expf(x - lse[B + j]) - (j == i ? 2.f : 0.f));
        }
        G[(size_t)i * ld + j] = g;
        if (Gh != nullptr) Gh[(size_t)i * ld + j] = __float2bfloat16_rn(g);
    }
}
__global__ void clip_finish_kernel(const float* __restrict__ scal, float inv_2B, float* __restrict__ loss) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0) loss[0] = scal[0] * inv_2B;
}

static int loss_setup(LossWS& w, int64_t batch, void* workspace, size_t workspace_bytes, const char* who) {
    TEAM_REQUIRE(batch >= 1 && batch <= LOSS_MAX_BATCH, "%s: batch %lld out of [1, %lld]", who, (long long)batch, (long long)LOSS_MAX_BATCH);
    loss_plan(batch, workspace, &w);
    if (workspace == nullptr || workspace_bytes < w.total) {
        set_error("%s: workspace %zu < %zu bytes", who, workspace_bytes, w.total);
        return TEAM_EWORKSPACE;
    }
    TEAM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
    return TEAM_OK;
}

}  // namespace team

using namespace team;

extern "C" size_t team_loss_workspace_bytes(int64_t batch) {
    if (batch < 1 || batch > LOSS_MAX_BATCH) return 0;
    LossWS w;
    loss_plan(batch, nullptr, &w);
    return w.total;
}

extern "C" int team_unicl_loss(int mode, const float* image, const float* text, const float* state, const int64_t* labels,
                               int64_t batch, float temperature, float grad_scale, float* losses, float* g_image,
                               float* g_text, float* g_state, void* workspace, size_t workspace_bytes, void* stream) {
    TEAM_REQUIRE(mode == TEAM_MODE_F32 || mode == TEAM_MODE_BF16, "unicl_loss: bad mode %d", mode);
    TEAM_REQUIRE(image && text && state && labels && losses && g_image && g_text && g_state, "unicl_loss: null pointer");
    TEAM_REQUIRE(temperature > 0.f, "unicl_loss: temperature must be positive");
    LossWS w;
    int rc = loss_setup(w, batch, workspace, workspace_bytes, "unicl_loss");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int B = (int)batch, ld = (int)w.G.ld;
    const bool bf = mode == TEAM_MODE_BF16;
    const float inv_tau = 1.0f / temperature;
    TEAM_LAUNCH(loss_normalize_kernel, (3 * B + 7) / 8, 256, 0, st, batch, image, text, state, w.X[0].f, w.X[1].f, w.X[2].f,
                bf ? w.X[0].h : nullptr, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr, w.inv[0], w.inv[1], w.inv[2], 3, 1);
    {   // sim = Xi Xi^T
        LSeg s{false, false, D, w.X[0], w.X[0]};
        if ((rc = loss_gemm(st, mode, w, B, B, w.sim, ld, &s, 1))) return rc;
    }
    TEAM_LAUNCH(unicl_cat_stats_kernel, (B + 7) / 8, 256, 0, st, B, ld, w.sim, labels, inv_tau, w.rmax, w.rpos, w.rall, w.rloss);
    TEAM_LAUNCH(loss_fold2_kernel, 1, 1024, 0, st, B, w.rmax + B, w.rloss, w.scal);                 // valid count, category sum
    TEAM_LAUNCH(unicl_cat_grad_kernel, (B + 7) / 8, 256, 0, st, B, ld, w.sim, labels, inv_tau, 0.5f * grad_scale, w.rmax, w.rpos, w.rall, w.scal, w.G.f, bf ? w.G.h : nullptr);
    {   // dcat = G Xi + G^T Xi
        LSeg s[2] = {{false, true, B, w.G, w.X[0]}, {true, true, B, w.G, w.X[0]}};
        if ((rc = loss_gemm(st, mode, w, B, D, w.dcat, D, s, 2))) return rc;
    }
    TEAM_LAUNCH(unicl_instance_kernel, (B + 7) / 8, 256, 0, st, B, w.X[0].f, w.X[1].f, w.X[2].f, w.inv[0], w.inv[1], w.inv[2], w.dcat, inv_tau, grad_scale / (3.0f * (float)B), g_image, g_text, g_state, w.rloss + B);
    TEAM_LAUNCH(loss_fold2_kernel, 1, 1024, 0, st, B, w.rloss + B, (const float*)nullptr, w.scal + 2);
    TEAM_LAUNCH(unicl_finish_kernel, 1, 32, 0, st, w.scal, 1.0f / (3.0f * (float)B), losses);
    return TEAM_OK;
}

extern "C" int team_clip_loss(int mode, const float* image, const float* text, int64_t batch, float logit_scale,
                              float grad_scale, float* loss, float* g_image, float* g_text, void* workspace,
                              size_t workspace_bytes, void* stream) {
    TEAM_REQUIRE(mode == TEAM_MODE_F32 || mode == TEAM_MODE_BF16, "clip_loss: bad mode %d", mode);
    TEAM_REQUIRE(image && text && loss && g_image && g_text, "clip_loss: null pointer");
    LossWS w;
    int rc = loss_setup(w, batch, workspace, workspace_bytes, "clip_loss");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int B = (int)batch, ld = (int)w.G.ld;
    const bool bf = mode == TEAM_MODE_BF16;
    // staged copies (bf16 shadows for the tensor-core path); ClipLoss does not normalise (the caller did)
    TEAM_LAUNCH(loss_normalize_kernel, (2 * B + 7) / 8, 256, 0, st, batch, image, text, (const float*)nullptr, w.X[0].f, w.X[1].f, (float*)nullptr,
                bf ? w.X[0].h : nullptr, bf ? w.X[1].h : nullptr, (__nv_bfloat16*)nullptr, (float*)nullptr, (float*)nullptr, (float*)nullptr, 2, 0);
    {
        LSeg s{false, false, D, w.X[0], w.X[1]};
        if ((rc = loss_gemm(st, mode, w, B, B, w.sim, ld, &s, 1))) return rc;                    // L = I T^T
        LSeg t{false, false, D, w.X[1], w.X[0]};
        if ((rc = loss_gemm(st, mode, w, B, B, w.simT, ld, &t, 1))) return rc;                   // L^T = T I^T
    }
    TEAM_LAUNCH(clip_stats_kernel, (2 * B + 7) / 8, 256, 0, st, B, ld, w.sim, w.simT, logit_scale, w.rmax, w.rloss);
    TEAM_LAUNCH(loss_fold2_kernel, 1, 1024, 0, st, 2 * B, w.rloss, (const float*)nullptr, w.scal);
    TEAM_LAUNCH(clip_grad_kernel, (B + 7) / 8, 256, 0, st, B, ld, w.sim, logit_scale, logit_scale * grad_scale / (2.0f * (float)B), w.rmax, w.G.f, bf ? w.G.h : nullptr);
    {
        LSeg s{false, true, B, w.G, w.X[1]};
        if ((rc = loss_gemm(st, mode, w, B, D, g_image, D, &s, 1))) return rc;                   // dI = G T
        LSeg t{true, true, B, w.G, w.X[0]};
        if ((rc = loss_gemm(st, mode, w, B, D, g_text, D, &t, 1))) return rc;                    // dT = G^T I
    }
    TEAM_LAUNCH(clip_finish_kernel, 1, 32, 0, st, w.scal, 1.0f / (2.0f * (float)B), loss);
    return TEAM_OK;
}
