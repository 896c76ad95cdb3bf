// Table-query rows (prototype outputs + state output of every sample), third generation: GRAM formulation.
//
// For sample b and table-query row k (k < C: prototype row k, k = C: the state-table row of the sample) the head forms
//     u = cw n + s + ai vi + at vt + as vs         n = NF row, s = S row + b_fc, vi / vt = own VF rows, vs = VF of the state row
//     xhat = LayerNorm-normalised u,   out_proto[b] = gamma (1/C) sum_{k<C} xhat_k + beta,   out_state[b] = gamma xhat_C + beta
// (convs/projections.py:84-87 through the factorisation of DESIGN.md section 3).  The second generation
// (head_table_kernels.cuh) evaluated every (sample, row) as a 512-wide vector with serial warp reductions for its
// LayerNorm statistics: ~330 warp instructions per row on a dependent chain.  u, xhat and dY = dL/du are linear
// combinations of {n, s, vi, vt, vs, gg, 1}, so (tests/gram_table_rows_model.py is the fp64 specification):
//   * every scalar of a (sample, row) pair - mean, variance, m2, the score gradients dY.v - follows from dot products
//     of those vectors.  sample x table dots come from ONE tensor-core GEMM of the forward,
//         W0 = VFo [2B,512] x [VFs ; S_table + b_fc ; NF_table]^T      (GEMM wave 4),
//     whose extra B-operand rows (the table rows s and n = sum_j P[r][j] VFs_j) are written by table_prep_kernel;
//     table x table dots are computed once per step (table_gram_prep_kernel, beside GEMM wave 4), sample x sample
//     dots once per sample; lane k of the sample's warp then does the whole LayerNorm algebra of row k in scalars -
//     all rows of a sample in parallel, no dependent 512-wide reductions;
//   * every vector output is  sum_k (coefficient_k x table row k)  +  scalars x the sample's own vectors: a pure
//     shared-memory FMA stream without reductions.
// The backward needs gg . n_k and gg . s_k (gg = gamma .* cotangent): 2C dots per sample, computed in fp32 in the
// kernel and reduced by ONE transposing butterfly (31 shuffles per 32 dots, lane k ends up with the dots of row k).
// Outputs, scratch layout (A1 / A23 coefficient matrices, per-CTA partial record) and determinism (fixed-order folds,
// no atomics) are those of the second generation, so everything downstream is unchanged.
// Needs C + 1 <= 32 and Rt <= 32 (lane = row) and the tables in shared memory; other heads use the older kernels.
#pragma once
#include "head_table_kernels.cuh"

namespace team {

constexpr int TG_MAXW = 16;         // warps per CTA at large batches (8 when the batch would not fill the SMs otherwise)
constexpr int TG_NRS = 8;           // floats per (sample, row) handed from the forward to the backward: rstd, mean, u.{n,s,vi,vt,vs}
constexpr int TG_RP = 32;           // rows reserved per table (s rows, n rows) behind the VF rows of the step (>= Rt)

// step-level dot products: [nn | ns | ss | sumN | sumS][32]  nvs[32][10]  svs[32][10]  vv[16]  sumV[16]
constexpr int TG_GT_NN = 0, TG_GT_NS = 32, TG_GT_SS = 64, TG_GT_SUMN = 96, TG_GT_SUMS = 128, TG_GT_NVS = 160,
              TG_GT_SVS = 480, TG_GT_VV = 800, TG_GT_SUMV = 816, TG_GT_LEN = 832;

// W0 [2B][ldw]: columns j < Nsp: VFo . VFs_j;  Nsp + tr: VFo . s_tr;  Nsp + 32 + tr: VFo . n_tr
__host__ __device__ inline int tg_ldw(const HeadDims& d) { return d.Nsp + 2 * TG_RP; }
__host__ __device__ inline size_t tg_fwd_smem_floats(const HeadDims& d) { return (size_t)2 * d.Rt * D + 2 * D + TG_GT_LEN; }
__host__ __device__ inline size_t tg_bwd_smem_floats(const HeadDims& d, int nw) {
    return (size_t)2 * d.Rt * D + D + TG_GT_LEN + (size_t)nw * D + (size_t)10 * D + (size_t)d.Rt * TQ_NSC + (size_t)nw * 32 * TB_NRS + 16;
}
__host__ __device__ inline bool table_gram_supported(const HeadDims& d) {
    return table2_supported(d) && tg_bwd_smem_floats(d, TG_MAXW) * sizeof(float) <= 227 * 1024;
}
__host__ __device__ inline int tg_warps(const HeadDims& d) { return (d.B + TG_MAXW - 1) / TG_MAXW >= NUM_SMS ? TG_MAXW : 8; }

// ------------------------------------------------------------------ step-level dot products (one warp per table row / state)
// X = the table rows behind the VF rows of the step: s rows X[0..32), n rows X[32..64) (table_prep_kernel); vst = VF rows
// of the ten states.  Runs beside GEMM wave 4; gt is read by the forward and the backward kernel.
__global__ void __launch_bounds__(256)
table_gram_prep_kernel(int Rt, const float* __restrict__ X, const float* __restrict__ vst, float* __restrict__ gt) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int tr = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tr >= Rt + 10) return;
    if (tr < Rt) {
        float4 n[4], s[4];
        ld_row(X + (size_t)(TG_RP + tr) * D, lane, n);
        ld_row(X + (size_t)tr * D, lane, s);
        float v0 = dot_part(n, n), v1 = dot_part(n, s), v2 = dot_part(s, s), v3 = sum_part(n), v4 = sum_part(s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o); v3 += __shfl_xor_sync(0xffffffffu, v3, o);
            v4 += __shfl_xor_sync(0xffffffffu, v4, o);
        }
        if (lane == 0) {
            gt[TG_GT_NN + tr] = v0; gt[TG_GT_NS + tr] = v1; gt[TG_GT_SS + tr] = v2;
            gt[TG_GT_SUMN + tr] = v3; gt[TG_GT_SUMS + tr] = v4;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {                 // five states at a time: ten interleaved reductions
            float a[5], c[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float4 v[4];
                ld_row(vst + (size_t)(5 * h + q) * D, lane, v);
                a[q] = dot_part(n, v); c[q] = dot_part(s, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 5; ++q) { a[q] += __shfl_xor_sync(0xffffffffu, a[q], o); c[q] += __shfl_xor_sync(0xffffffffu, c[q], o); }
            }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 5; ++q) { gt[TG_GT_NVS + tr * 10 + 5 * h + q] = a[q]; gt[TG_GT_SVS + tr * 10 + 5 * h + q] = c[q]; }
            }
        }
    } else {
        const int st = tr - Rt;
        float4 v[4];
        ld_row(vst + (size_t)st * D, lane, v);
        float v0 = dot_part(v, v), v1 = sum_part(v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o); }
        if (lane == 0) { gt[TG_GT_VV + st] = v0; gt[TG_GT_SUMV + st] = v1; }
    }
}

// table rows + step-level dots into shared memory (the caller synchronises)
__device__ __forceinline__ void tg_load(const HeadDims& d, const float* __restrict__ X, const float* __restrict__ gtg,
                                        float* tabN, float* tabS, float* gt) {
    const int n4 = d.Rt * (D / 4);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        reinterpret_cast<float4*>(tabS)[i] = reinterpret_cast<const float4*>(X)[i];
        reinterpret_cast<float4*>(tabN)[i] = reinterpret_cast<const float4*>(X + (size_t)TG_RP * D)[i];
    }
    for (int i = threadIdx.x; i < TG_GT_LEN; i += blockDim.x) gt[i] = gtg[i];
}

// ------------------------------------------------------------------ forward
// dynamic smem: tabN[Rt][D] | tabS[Rt][D] | gam[D] | bet[D] | gt[TG_GT_LEN]
// Saved for the backward: xs[b] = sum_{k<C} xhat_k, xst[b] = xhat_C, RS[b][TG_NRS][32] (lane = row).
__global__ void __launch_bounds__(TG_MAXW * 32, 1)
table_gram_fwd_kernel(HeadDims d, const float* __restrict__ SK, const float* __restrict__ TT,
                      const float* __restrict__ mt, const float* __restrict__ Zt, const float* __restrict__ X,
                      const float* __restrict__ gtg, const float* __restrict__ W0,
                      const float* __restrict__ VFo, const float* __restrict__ VFs,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      const int64_t* __restrict__ state_ids, float* __restrict__ out_proto,
                      float* __restrict__ out_state, float* __restrict__ xs, float* __restrict__ xst,
                      float* __restrict__ RS) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float tg_smem[];
    const int ldw = tg_ldw(d);
    float* tabN = tg_smem;
    float* tabS = tabN + (size_t)d.Rt * D;
    float* gam = tabS + (size_t)d.Rt * D;
    float* bet = gam + D;
    float* gt = bet + D;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int i = tid; i < D; i += blockDim.x) { gam[i] = gamma[i]; bet[i] = beta[i]; }
    tg_load(d, X, gtg, tabN, tabS, gt);
    __syncthreads();
    const float invC = 1.0f / (float)d.C;
    for (int b = blockIdx.x * nw + warp; b < d.B; b += gridDim.x * nw) {
        const int sid = clamp_state(state_ids[b]);
        const int srow = d.M + sid;
        const float* w0 = W0 + (size_t)b * ldw;
        const float* w1 = W0 + (size_t)(d.B + b) * ldw;
        // ---- scalar inputs of lane k first (global loads in flight under the sample x sample dots)
        TableRowW rw;
        rw.c_w = rw.a_i = rw.a_t = rw.a_s = 0.f; rw.r = 0;
        float nvi = 0.f, nvt = 0.f, svi = 0.f, svt = 0.f;
        const int tr = lane < d.C ? lane : d.C + sid;
        if (lane <= d.C) {
            svi = w0[d.Nsp + tr]; svt = w1[d.Nsp + tr];
            nvi = w0[d.Nsp + TG_RP + tr]; nvt = w1[d.Nsp + TG_RP + tr];
            rw = table_row_weights(d, b, lane, srow, SK, TT, mt, Zt);
        }
        const float vsvi = w0[srow], vsvt = w1[srow];
        float4 vi[4], vt[4];
        ld_row(VFo + (size_t)b * D, lane, vi);
        ld_row(VFo + (size_t)(d.B + b) * D, lane, vt);
        float vivi = dot_part(vi, vi), vivt = dot_part(vi, vt), vtvt = dot_part(vt, vt), smi = sum_part(vi), smt = sum_part(vt);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vivi += __shfl_xor_sync(0xffffffffu, vivi, o); vivt += __shfl_xor_sync(0xffffffffu, vivt, o);
            vtvt += __shfl_xor_sync(0xffffffffu, vtvt, o); smi += __shfl_xor_sync(0xffffffffu, smi, o);
            smt += __shfl_xor_sync(0xffffffffu, smt, o);
        }
        // ---- scalar phase: lane k owns table-query row k
        float cn = 0.f, cs = 0.f, ci = 0.f, ct = 0.f, cv = 0.f, c0 = 0.f;       // coefficients of xhat_k (rows k < C)
        float x_r = 0.f, x_cw = 0.f, x_ai = 0.f, x_at = 0.f, x_as = 0.f, x_mean = 0.f;   // the state row (lane C)
        if (lane <= d.C) {
            const float nn = gt[TG_GT_NN + tr], ns = gt[TG_GT_NS + tr], ss = gt[TG_GT_SS + tr];
            const float nvs = gt[TG_GT_NVS + tr * 10 + sid], svs = gt[TG_GT_SVS + tr * 10 + sid], vv = gt[TG_GT_VV + sid];
            const float cw = rw.c_w, ai = rw.a_i, at = rw.a_t, as = rw.a_s;
            const float uvn = fmaf(cw, nn, ns) + fmaf(ai, nvi, fmaf(at, nvt, as * nvs));
            const float uvs = fmaf(cw, ns, ss) + fmaf(ai, svi, fmaf(at, svt, as * svs));
            const float uvi = fmaf(cw, nvi, svi) + fmaf(ai, vivi, fmaf(at, vivt, as * vsvi));
            const float uvt = fmaf(cw, nvt, svt) + fmaf(ai, vivt, fmaf(at, vtvt, as * vsvt));
            const float uvv = fmaf(cw, nvs, svs) + fmaf(ai, vsvi, fmaf(at, vsvt, as * vv));
            const float uu = fmaf(cw, uvn, uvs) + fmaf(ai, uvi, fmaf(at, uvt, as * uvv));
            const float su = fmaf(cw, gt[TG_GT_SUMN + tr], gt[TG_GT_SUMS + tr]) + fmaf(ai, smi, fmaf(at, smt, as * gt[TG_GT_SUMV + sid]));
            const float mean = su * (1.0f / D);
            const float var = fmaxf(uu * (1.0f / D) - mean * mean, 0.f);
            const float rstd = 1.0f / sqrtf(var + LN_EPS);
            float* rs = RS + (size_t)b * TG_NRS * 32 + lane;
            rs[0] = rstd; rs[32] = mean; rs[64] = uvn; rs[96] = uvs; rs[128] = uvi; rs[160] = uvt; rs[192] = uvv;
            if (lane < d.C) {
                cn = rstd * cw; cs = rstd; ci = rstd * ai; ct = rstd * at; cv = rstd * as; c0 = rstd * mean;
            } else {
                x_r = rstd; x_cw = cw; x_ai = ai; x_at = at; x_as = as; x_mean = mean;
            }
        }
        // ---- vector phase: sum_k (cn_k n_k + cs_k s_k) + own terms; two rows per iteration (loads in flight)
        float4 acc[4];
        zero_row(acc);
        int k = 0;
        for (; k + 1 < d.C; k += 2) {
            const float a0 = __shfl_sync(0xffffffffu, cn, k), s0 = __shfl_sync(0xffffffffu, cs, k);
            const float a1 = __shfl_sync(0xffffffffu, cn, k + 1), s1 = __shfl_sync(0xffffffffu, cs, k + 1);
            float4 u0[4], u1[4], u2[4], u3[4];
            ld_row(tabN + (size_t)k * D, lane, u0);
            ld_row(tabS + (size_t)k * D, lane, u1);
            ld_row(tabN + (size_t)(k + 1) * D, lane, u2);
            ld_row(tabS + (size_t)(k + 1) * D, lane, u3);
            axpy_row(acc, a0, u0); axpy_row(acc, s0, u1); axpy_row(acc, a1, u2); axpy_row(acc, s1, u3);
        }
        for (; k < d.C; ++k) {
            const float a = __shfl_sync(0xffffffffu, cn, k), s_ = __shfl_sync(0xffffffffu, cs, k);
            float4 u[4];
            ld_row(tabN + (size_t)k * D, lane, u);
            axpy_row(acc, a, u);
            ld_row(tabS + (size_t)k * D, lane, u);
            axpy_row(acc, s_, u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ci += __shfl_xor_sync(0xffffffffu, ci, o); ct += __shfl_xor_sync(0xffffffffu, ct, o);
            cv += __shfl_xor_sync(0xffffffffu, cv, o); c0 += __shfl_xor_sync(0xffffffffu, c0, o);
        }
        float4 vs[4], g[4], be[4];
        ld_row(VFs + (size_t)srow * D, lane, vs);
        ld_row(gam, lane, g);
        ld_row(bet, lane, be);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = add4s(-c0, fma4s(ci, vi[i], fma4s(ct, vt[i], fma4s(cv, vs[i], acc[i]))));
        st_row(xs + (size_t)b * D, lane, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fma4(mul4s(invC, acc[i]), g[i], be[i]);
        st_row(out_proto + (size_t)b * D, lane, acc);
        {   // the state row: one row per sample, directly
            x_r = __shfl_sync(0xffffffffu, x_r, d.C); x_cw = __shfl_sync(0xffffffffu, x_cw, d.C);
            x_ai = __shfl_sync(0xffffffffu, x_ai, d.C); x_at = __shfl_sync(0xffffffffu, x_at, d.C);
            x_as = __shfl_sync(0xffffffffu, x_as, d.C); x_mean = __shfl_sync(0xffffffffu, x_mean, d.C);
            const int trS = d.C + sid;
            float4 n[4], s[4];
            ld_row(tabN + (size_t)trS * D, lane, n);
            ld_row(tabS + (size_t)trS * D, lane, s);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 u = fma4s(x_cw, n[i], fma4s(x_ai, vi[i], fma4s(x_at, vt[i], fma4s(x_as, vs[i], s[i]))));
                n[i] = mul4s(x_r, add4s(-x_mean, u));
            }
            st_row(xst + (size_t)b * D, lane, n);
#pragma unroll
            for (int i = 0; i < 4; ++i) n[i] = fma4(n[i], g[i], be[i]);
            st_row(out_state + (size_t)b * D, lane, n);
        }
    }
}

// ------------------------------------------------------------------ backward: LayerNorm gamma / beta gradients of the table rows
// dgamma += gp/C .* xs + gs .* xst, dbeta += gp + gs over the batch: thread-owned float2 columns, samples in index order
// (deterministic).  Same grid as table_gram_bwd_kernel: CTA c fills the dgam / dbet fields of partial record c.  Depends
// only on the forward's xs / xst and the cotangents, so it runs on the side lane beside the table-row kernel.
__global__ void __launch_bounds__(256)
table_dgamma_kernel(HeadDims d, const float* __restrict__ g_proto, const float* __restrict__ g_state,
                    const float* __restrict__ xs, const float* __restrict__ xst, float* __restrict__ partials) {
    pdl_trigger();
    pdl_wait();
    const TabOff off = tab_offsets(d);
    const int tid = threadIdx.x;
    const float invC = d.C > 1 ? 1.0f / (float)d.C : 1.0f;
    float2 dg = make_float2(0.f, 0.f), db = dg;
    const float fc = (float)d.C;
    for (int b0 = blockIdx.x; b0 < d.B; b0 += 4 * gridDim.x) {          // four samples per iteration: 16 loads in flight
        float2 gp[4], gs[4], x0[4], x1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b = b0 + q * gridDim.x;
            gp[q] = gs[q] = x0[q] = x1[q] = make_float2(0.f, 0.f);
            if (b < d.B) {
                if (g_proto != nullptr) gp[q] = reinterpret_cast<const float2*>(g_proto + (size_t)b * D)[tid];
                gs[q] = reinterpret_cast<const float2*>(g_state + (size_t)b * D)[tid];
                x0[q] = reinterpret_cast<const float2*>(xs + (size_t)b * D)[tid];
                x1[q] = reinterpret_cast<const float2*>(xst + (size_t)b * D)[tid];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                   // samples in index order (deterministic)
            dg.x += invC * gp[q].x * x0[q].x + gs[q].x * x1[q].x; dg.y += invC * gp[q].y * x0[q].y + gs[q].y * x1[q].y;
            db.x += fc * (invC * gp[q].x) + gs[q].x; db.y += fc * (invC * gp[q].y) + gs[q].y;
        }
    }
    float* rec = partials + (size_t)blockIdx.x * off.len;
    reinterpret_cast<float2*>(rec + off.dgam)[tid] = dg;
    reinterpret_cast<float2*>(rec + off.dbet)[tid] = db;
}

// ------------------------------------------------------------------ backward
// dynamic smem: tabN[Rt][D] | tabS[Rt][D] | gam[D] | gt[TG_GT_LEN] | slots[nw][D] (sum_k a_s dY) | dvfst[10][D]
//               | scal[Rt][TQ_NSC] | rsl[nw][32][TB_NRS] | sidw[nw]
// Per-CTA partial record (except its dgam / dbet fields: table_dgamma_kernel) and the scalar outputs: exactly those of
// table_rows_bwd2_kernel.  Register budget 128 (16 warps per SM): vectors are re-read where they are needed instead of
// being held across phases.
__global__ void __launch_bounds__(TG_MAXW * 32, 1)
table_gram_bwd_kernel(HeadDims d, const float* __restrict__ SK, const float* __restrict__ TT,
                      const float* __restrict__ mt, const float* __restrict__ Zt, const float* __restrict__ X,
                      const float* __restrict__ gtg, const float* __restrict__ VFo, const float* __restrict__ VFs,
                      const float* __restrict__ gamma,
                      const int64_t* __restrict__ state_ids, const float* __restrict__ g_proto,
                      const float* __restrict__ g_state, const float* __restrict__ RS,
                      float* __restrict__ dSK, __nv_bfloat16* __restrict__ dSKh,
                      float* __restrict__ dVFo, float* __restrict__ GG, __nv_bfloat16* __restrict__ GGh,
                      float* __restrict__ A1, __nv_bfloat16* __restrict__ A1h, float* __restrict__ A23,
                      __nv_bfloat16* __restrict__ A23h, int ldA, float* __restrict__ partials) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float tg_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float* tabN = tg_smem;
    float* tabS = tabN + (size_t)d.Rt * D;
    float* gam = tabS + (size_t)d.Rt * D;
    float* gt = gam + D;
    float* slots = gt + TG_GT_LEN;                           // [nw][D]
    float* dvfst = slots + (size_t)nw * D;                   // [10][D]
    float* scal = dvfst + 10 * D;                            // [Rt][TQ_NSC]
    float* rsl = scal + (size_t)d.Rt * TQ_NSC;               // [nw][32][TB_NRS]
    int* sidw = reinterpret_cast<int*>(rsl + (size_t)nw * 32 * TB_NRS);
    const TabOff off = tab_offsets(d);
    const int gcol = ldA / 2;
    for (int i = tid; i < 10 * D; i += blockDim.x) dvfst[i] = 0.f;
    for (int i = tid; i < d.Rt * TQ_NSC; i += blockDim.x) scal[i] = 0.f;
    for (int i = tid; i < D; i += blockDim.x) gam[i] = gamma[i];
    tg_load(d, X, gtg, tabN, tabS, gt);
    const float invC = d.C > 1 ? 1.0f / (float)d.C : 1.0f;
    const __nv_bfloat16 hz = __float2bfloat16_rn(0.f);
    const bool has_p = g_proto != nullptr;
    __syncthreads();
    const int ngroups = (d.B + nw - 1) / nw;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int b = grp * nw + warp;
        if (b < d.B) {
            const int sid = clamp_state(state_ids[b]);
            const int srow = d.M + sid;
            const int trS = d.C + sid;
            // lane k's scalar inputs first: their global loads are in flight under the vector work below
            TableRowW mine;
            mine.c_w = mine.a_i = mine.a_t = mine.a_s = 0.f; mine.r = 0;
            float f_rstd = 0.f, f_mean = 0.f, uvn = 0.f, uvi = 0.f, uvt = 0.f, uvv = 0.f;
            if (lane <= d.C) {
                mine = table_row_weights(d, b, lane, srow, SK, TT, mt, Zt);
                const float* rs = RS + (size_t)b * TG_NRS * 32 + lane;
                f_rstd = rs[0]; f_mean = rs[32]; uvn = rs[64]; uvi = rs[128]; uvt = rs[160]; uvv = rs[192];
            }
            for (int i = lane; i < d.Nsp; i += 32) {
                dSK[(size_t)b * d.Nsp + i] = 0.f;
                dSK[(size_t)(d.B + b) * d.Nsp + i] = 0.f;
                if (dSKh != nullptr) { dSKh[(size_t)b * d.Nsp + i] = hz; dSKh[(size_t)(d.B + b) * d.Nsp + i] = hz; }
            }
            for (int i = lane; i < ldA; i += 32) {
                A1[(size_t)b * ldA + i] = 0.f; A1[(size_t)(d.B + b) * ldA + i] = 0.f;
                A23[(size_t)b * ldA + i] = 0.f; A23[(size_t)(d.B + b) * ldA + i] = 0.f;
                if (A1h != nullptr) {
                    A1h[(size_t)b * ldA + i] = hz; A1h[(size_t)(d.B + b) * ldA + i] = hz;
                    A23h[(size_t)b * ldA + i] = hz; A23h[(size_t)(d.B + b) * ldA + i] = hz;
                }
            }
            // ---- cotangents: GG rows (cotangent .* gamma, rows of the coefficient GEMM); all dots that involve gg_s
            float r_[12];
            float4 ggp[4];
            {
                float4 g4[4], ggs[4], t[4];
                ld_row(gam, lane, g4);
                ld_row(g_state + (size_t)b * D, lane, ggs);
                if (has_p) ld_row(g_proto + (size_t)b * D, lane, ggp); else zero_row(ggp);
#pragma unroll
                for (int i = 0; i < 4; ++i) { ggs[i] = mul4(ggs[i], g4[i]); ggp[i] = mul4(mul4s(invC, ggp[i]), g4[i]); }
                st_row(GG + (size_t)b * D, lane, ggp);
                st_row(GG + (size_t)(d.B + b) * D, lane, ggs);
                if (GGh != nullptr) { st_row_h(GGh + (size_t)b * D, lane, ggp); st_row_h(GGh + (size_t)(d.B + b) * D, lane, ggs); }
                r_[0] = sum_part(ggp); r_[1] = sum_part(ggs);
                ld_row(VFo + (size_t)b * D, lane, t);
                r_[2] = dot_part(ggp, t); r_[4] = dot_part(ggs, t); r_[6] = sum_part(t);
                ld_row(VFo + (size_t)(d.B + b) * D, lane, t);
                r_[3] = dot_part(ggp, t); r_[5] = dot_part(ggs, t); r_[7] = sum_part(t);
                ld_row(VFs + (size_t)srow * D, lane, t);
                r_[8] = dot_part(ggp, t); r_[9] = dot_part(ggs, t);
                ld_row(tabN + (size_t)trS * D, lane, t);
                r_[10] = dot_part(ggs, t);
                ld_row(tabS + (size_t)trS * D, lane, t);
                r_[11] = dot_part(ggs, t);
            }
            __syncwarp();                                     // the zero fill is ordered before the per-row entries below
            // ---- gg_p . table rows: 2C dots, transposing butterfly -> lane k holds gg_p . n_k and gg_p . s_k
            // (one table at a time: 32 partial dots live instead of 64 - register budget)
            float gpn = 0.f, gps = 0.f;
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const float* tab = which == 0 ? tabN : tabS;
                float pp[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    pp[k] = 0.f;
                    if (has_p && k < d.C) {
                        float4 u[4];
                        ld_row(tab + (size_t)k * D, lane, u);
                        pp[k] = dot_part(ggp, u);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const bool up = (lane & o) != 0;
#pragma unroll
                    for (int i = 0; i < o; ++i) {
                        const float kn = up ? pp[i + o] : pp[i], sn = up ? pp[i] : pp[i + o];
                        pp[i] = kn + __shfl_xor_sync(0xffffffffu, sn, o);
                    }
                }
                if (which == 0) gpn = pp[0]; else gps = pp[0];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 12; ++q) r_[q] += __shfl_xor_sync(0xffffffffu, r_[q], o);
            }
            // ---- scalar phase: lane k owns table-query row k
            float k_alpha = 0.f, k_beta = 0.f, k_mean = 0.f, k_m1 = 0.f, k_yy = 0.f, k_i = 0.f, k_t = 0.f, k_s = 0.f, k_de = 0.f;
            if (lane <= d.C && (has_p || lane == d.C)) {
                const bool is_p = lane < d.C;
                const int tr = is_p ? lane : trS;
                const float g_n = is_p ? gpn : r_[10], g_s = is_p ? gps : r_[11];
                const float g_i = is_p ? r_[2] : r_[4], g_t = is_p ? r_[3] : r_[5], g_v = is_p ? r_[8] : r_[9];
                const float sg = is_p ? r_[0] : r_[1];
                const float cw = mine.c_w, ai = mine.a_i, at = mine.a_t, as = mine.a_s;
                const float m1 = sg * (1.0f / D);
                const float gu = fmaf(cw, g_n, g_s) + fmaf(ai, g_i, fmaf(at, g_t, as * g_v));       // gg . u
                const float m2 = f_rstd * (gu - f_mean * sg) * (1.0f / D);
                const float al = f_rstd, be = f_rstd * f_rstd * m2;
                const float de = fmaf(be, f_mean, -f_rstd * m1);
                const float dYn = fmaf(al, g_n, fmaf(-be, uvn, de * gt[TG_GT_SUMN + tr]));
                const float dYi = fmaf(al, g_i, fmaf(-be, uvi, de * r_[6]));
                const float dYt = fmaf(al, g_t, fmaf(-be, uvt, de * r_[7]));
                const float dYv = fmaf(al, g_v, fmaf(-be, uvv, de * gt[TG_GT_SUMV + sid]));
                k_yy = fmaf(cw, dYn, fmaf(ai, dYi, fmaf(at, dYt, as * dYv)));                        // dY . ybar, ybar = u - s
                k_i = dYi; k_t = dYt; k_s = dYv;
                k_alpha = al; k_beta = be; k_mean = f_mean; k_m1 = m1; k_de = de;
            }
            // ---- vector phase: sum_k a_q dY_k for q = image / text / state key
            //      dY_k = al_k gg - be_k u_k + de_k,  u_k = cw_k n_k + s_k + ai_k vi + at_k vt + as_k vs
            float4 acc_i[4], acc_t[4], acc_s[4];
            zero_row(acc_i); zero_row(acc_t); zero_row(acc_s);
            const float wi = mine.a_i * k_beta, wt = mine.a_t * k_beta, ws = mine.a_s * k_beta;
            for (int k = has_p ? 0 : d.C; k <= d.C; ++k) {
                const float b_i = -__shfl_sync(0xffffffffu, wi, k), b_t = -__shfl_sync(0xffffffffu, wt, k);
                const float b_s = -__shfl_sync(0xffffffffu, ws, k), cw = __shfl_sync(0xffffffffu, mine.c_w, k);
                const int tr = k < d.C ? k : trS;
                float4 n[4], s[4];
                ld_row(tabN + (size_t)tr * D, lane, n);
                ld_row(tabS + (size_t)tr * D, lane, s);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 u = fma4s(cw, n[i], s[i]);
                    acc_i[i] = fma4s(b_i, u, acc_i[i]); acc_t[i] = fma4s(b_t, u, acc_t[i]); acc_s[i] = fma4s(b_s, u, acc_s[i]);
                }
            }
            {   // own-vector terms: coefficient sums over the rows (15 interleaved reductions + 3 broadcasts from the state lane)
                const bool is_p = lane < d.C;
                float c_[15];
                const float a3[3] = {mine.a_i, mine.a_t, mine.a_s}, w3[3] = {wi, wt, ws};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    c_[5 * q + 0] = is_p ? a3[q] * k_alpha : 0.f;           // x gg_p
                    c_[5 * q + 1] = w3[q] * mine.a_i;                       // x vi (subtracted)
                    c_[5 * q + 2] = w3[q] * mine.a_t;
                    c_[5 * q + 3] = w3[q] * mine.a_s;
                    c_[5 * q + 4] = a3[q] * k_de;                           // x 1
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int q = 0; q < 15; ++q) c_[q] += __shfl_xor_sync(0xffffffffu, c_[q], o);
                }
                const float al_s = __shfl_sync(0xffffffffu, k_alpha, d.C);
                const float ks_i = __shfl_sync(0xffffffffu, mine.a_i, d.C) * al_s, ks_t = __shfl_sync(0xffffffffu, mine.a_t, d.C) * al_s;
                const float ks_s = __shfl_sync(0xffffffffu, mine.a_s, d.C) * al_s;
                float4 t[4];                                                  // one vector at a time (register budget)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc_i[i] = add4s(c_[4], fma4s(c_[0], ggp[i], acc_i[i]));
                    acc_t[i] = add4s(c_[9], fma4s(c_[5], ggp[i], acc_t[i]));
                    acc_s[i] = add4s(c_[14], fma4s(c_[10], ggp[i], acc_s[i]));
                }
                ld_row(GG + (size_t)(d.B + b) * D, lane, t);                  // gg_s: written above by these same lanes
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc_i[i] = fma4s(ks_i, t[i], acc_i[i]); acc_t[i] = fma4s(ks_t, t[i], acc_t[i]); acc_s[i] = fma4s(ks_s, t[i], acc_s[i]); }
                ld_row(VFo + (size_t)b * D, lane, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc_i[i] = fma4s(-c_[1], t[i], acc_i[i]); acc_t[i] = fma4s(-c_[6], t[i], acc_t[i]); acc_s[i] = fma4s(-c_[11], t[i], acc_s[i]); }
                ld_row(VFo + (size_t)(d.B + b) * D, lane, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc_i[i] = fma4s(-c_[2], t[i], acc_i[i]); acc_t[i] = fma4s(-c_[7], t[i], acc_t[i]); acc_s[i] = fma4s(-c_[12], t[i], acc_s[i]); }
                ld_row(VFs + (size_t)srow * D, lane, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc_i[i] = fma4s(-c_[3], t[i], acc_i[i]); acc_t[i] = fma4s(-c_[8], t[i], acc_t[i]); acc_s[i] = fma4s(-c_[13], t[i], acc_s[i]); }
            }
            st_row(dVFo + (size_t)b * D, lane, acc_i);
            st_row(dVFo + (size_t)(d.B + b) * D, lane, acc_t);
            st_row(slots + (size_t)warp * D, lane, acc_s);
            if (lane <= d.C) {                               // scalar outputs, one row per lane (as table_rows_bwd2_kernel)
                const bool is_proto = lane < d.C;
                const int tr = is_proto ? lane : trS;
                const float cw = mine.c_w;
                const float v_i = mine.a_i * (k_i - k_yy) * INV_TAU, v_t = mine.a_t * (k_t - k_yy) * INV_TAU;
                dSK[(size_t)b * d.Nsp + mine.r] = v_i;
                dSK[(size_t)(d.B + b) * d.Nsp + mine.r] = v_t;
                if (dSKh != nullptr) {
                    dSKh[(size_t)b * d.Nsp + mine.r] = __float2bfloat16_rn(v_i);
                    dSKh[(size_t)(d.B + b) * d.Nsp + mine.r] = __float2bfloat16_rn(v_t);
                }
                const size_t ra = (size_t)(is_proto ? b : d.B + b) * ldA;        // GG row this query's cotangent lives in
                const size_t r0 = (size_t)b * ldA, r1 = (size_t)(d.B + b) * ldA;
                const float c23i = -k_beta * mine.a_i, c23t = -k_beta * mine.a_t;
                A1[ra + tr] = k_alpha; A1[ra + gcol + tr] = cw * k_alpha;
                A23[r0 + tr] = c23i; A23[r0 + gcol + tr] = cw * c23i;
                A23[r1 + tr] = c23t; A23[r1 + gcol + tr] = cw * c23t;
                if (A1h != nullptr) {
                    A1h[ra + tr] = __float2bfloat16_rn(k_alpha); A1h[ra + gcol + tr] = __float2bfloat16_rn(cw * k_alpha);
                    A23h[r0 + tr] = __float2bfloat16_rn(c23i); A23h[r0 + gcol + tr] = __float2bfloat16_rn(cw * c23i);
                    A23h[r1 + tr] = __float2bfloat16_rn(c23t); A23h[r1 + gcol + tr] = __float2bfloat16_rn(cw * c23t);
                }
                const float e0 = k_alpha * k_m1, e1 = k_beta * cw, e2 = k_beta, e3 = k_beta * k_mean, e4 = k_beta * mine.a_s;
                float* rs = rsl + ((size_t)warp * 32 + lane) * TB_NRS;
                rs[0] = e0; rs[1] = e1; rs[2] = e2; rs[3] = e3;
                rs[4] = cw * e0; rs[5] = cw * e1; rs[6] = cw * e2; rs[7] = cw * e3;
                rs[8] = cw * k_yy;
                rs[9] = mine.a_s * (k_s - k_yy) * INV_TAU; rs[10] = e4; rs[11] = cw * e4;
            }
            if (lane == 0) sidw[warp] = sid;
        } else if (lane == 0) {
            sidw[warp] = -1;
        }
        __syncthreads();
        // ---- fold of the round: thread-owned columns, warps in order
        for (int c = tid; c < D / 2; c += blockDim.x) {
            for (int w = 0; w < nw; ++w) {
                const int s = sidw[w];
                if (s < 0) continue;
                const float2 as = reinterpret_cast<const float2*>(slots + (size_t)w * D)[c];
                float2* dv = reinterpret_cast<float2*>(dvfst + (size_t)s * D) + c;
                float2 o = *dv;
                o.x += as.x; o.y += as.y;
                *dv = o;
            }
        }
        for (int t = tid; t < (d.C + 1) * TB_NRS; t += blockDim.x) {
            const int j = t / TB_NRS, kk = t - j * TB_NRS;
            for (int w = 0; w < nw; ++w) {
                const int s = sidw[w];
                if (s < 0) continue;
                const int tr = j < d.C ? j : d.C + s;
                const int idx = kk < 9 ? kk : 10 * (kk - 8) + s;
                scal[tr * TQ_NSC + idx] += rsl[((size_t)w * 32 + j) * TB_NRS + kk];
            }
        }
        __syncthreads();
    }
    float* rec = partials + (size_t)blockIdx.x * off.len;
    for (int i = tid; i < 10 * D; i += blockDim.x) rec[off.dvfst + i] = dvfst[i];
    for (int i = tid; i < d.Rt * TQ_NSC; i += blockDim.x) rec[off.scal + i] = scal[i];
    for (size_t i = off.scal + (size_t)d.Rt * TQ_NSC + tid; i < off.len; i += blockDim.x) rec[i] = 0.f;
}

}  // namespace team
