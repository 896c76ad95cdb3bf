// Backward kernels of the fusion head (hand-written VJP of head_fwd_kernels.cuh; replaces
// autograd through utils/inc_net.py:528-580 + convs/projections.py:64-87 at models/proof.py:444).
// All batch reductions (shared-row gradients, LayerNorm/bias gradients) are two-pass and
// fixed-order: per-CTA partial records, then reduce_partials_kernel -> bit-reproducible.
#pragma once
#include "head_fwd_kernels.cuh"

namespace team {

// offsets (floats) inside one table partial record: [R][G][dvfst][dgam][dbet][h][dtts][pad]
struct TabOff {
    size_t R, G, dvfst, dgam, dbet, h, dtts, len;
};
__host__ __device__ inline TabOff tab_offsets(const HeadDims& d) {
    TabOff o;
    o.R = 0;
    o.G = o.R + (size_t)d.Rt * D;
    o.dvfst = o.G + (size_t)d.Rt * D;
    o.dgam = o.dvfst + (size_t)10 * D;
    o.dbet = o.dgam + D;
    o.h = o.dbet + D;
    o.dtts = o.h + d.Rt;
    o.len = (o.dtts + (size_t)d.Rt * 10 + 3) / 4 * 4;
    return o;
}

// ------------------------------------------------------------------ table-query rows, backward
// dynamic smem: slot[3][TR_WARPS][D] | dvfst[10][D] | lnp[3][D] (gamma,beta,bfc) | hacc[Rt] | dtts[Rt][10]
__global__ void __launch_bounds__(TR_WARPS * 32)
table_rows_bwd_kernel(HeadDims d, const float* __restrict__ SK, const float* __restrict__ TT,
                      const float* __restrict__ mt, const float* __restrict__ Zt, const float* __restrict__ NFt,
                      const float* __restrict__ VFo, const float* __restrict__ VFs, const float* __restrict__ S,
                      const float* __restrict__ bfc, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const int64_t* __restrict__ state_ids, const float* __restrict__ g_proto,
                      const float* __restrict__ g_state, float* __restrict__ dSK, __nv_bfloat16* __restrict__ dSKh,
                      float* __restrict__ dVFo, float* __restrict__ partials) {
    extern __shared__ __align__(16) float tb_smem[];
    float* slot = tb_smem;                                    // [3][TR_WARPS][D]
    float* dvfst = slot + 3 * TR_WARPS * D;                   // [10][D]
    float* lnp = dvfst + 10 * D;                              // [3][D]
    float* hacc = lnp + 3 * D;                                // [Rt]
    float* dtts = hacc + d.Rt;                                // [Rt][10]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TabOff off = tab_offsets(d);
    float* rec = partials + (size_t)blockIdx.x * off.len;
    for (size_t i = tid; i < off.len; i += blockDim.x) rec[i] = 0.f;
    for (int i = tid; i < 10 * D; i += blockDim.x) dvfst[i] = 0.f;
    for (int i = tid; i < D; i += blockDim.x) { lnp[i] = gamma[i]; lnp[D + i] = beta[i]; lnp[2 * D + i] = bfc[i]; }
    for (int i = tid; i < d.Rt * 11; i += blockDim.x) hacc[i] = 0.f;        // hacc + dtts are contiguous
    __syncthreads();
    float4 dgam[4], dbet[4];
    zero_row(dgam); zero_row(dbet);
    const int rounds = (d.C + 1 + TR_WARPS - 1) / TR_WARPS;
    const float invC = d.C > 1 ? 1.0f / (float)d.C : 1.0f;
    for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
        const int sid = clamp_state(state_ids[b]);
        const int srow = d.M + sid;
        for (int i = tid; i < d.Nsp; i += blockDim.x) {
            dSK[(size_t)b * d.Nsp + i] = 0.f;
            dSK[(size_t)(d.B + b) * d.Nsp + i] = 0.f;
            if (dSKh != nullptr) {
                dSKh[(size_t)b * d.Nsp + i] = __float2bfloat16_rn(0.f);
                dSKh[(size_t)(d.B + b) * d.Nsp + i] = __float2bfloat16_rn(0.f);
            }
        }
        __syncthreads();
        float2 acc_i = make_float2(0.f, 0.f), acc_t = acc_i, acc_s = acc_i;   // columns 2*tid, 2*tid+1
        for (int rd = 0; rd < rounds; ++rd) {
            const int j = rd * TR_WARPS + warp;
            if (j <= d.C) {
                TableRowCtx cx;
                float4 ybar[4], u[4], t[4], xh[4], du[4];
                table_row_forward(d, b, j, srow, lane, SK, TT, mt, Zt, NFt, VFo, VFs, cx, ybar);
                ld_row(S + (size_t)cx.r * D, lane, u);
                add_row(u, ybar);
                ld_row(lnp + 2 * D, lane, t); add_row(u, t);
                float rstd;
                {   // LayerNorm forward (normalised values only)
                    const float mean = warp_sum(sum_part(u)) * (1.0f / D);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { xh[i].x = u[i].x - mean; xh[i].y = u[i].y - mean; xh[i].z = u[i].z - mean; xh[i].w = u[i].w - mean; }
                    const float var = warp_sum(dot_part(xh, xh)) * (1.0f / D);
                    rstd = 1.0f / sqrtf(var + LN_EPS);
                    scale_row(xh, rstd);
                }
                // cotangent of this row
                ld_row((j < d.C ? g_proto : g_state) + (size_t)b * D, lane, u);
                if (j < d.C) scale_row(u, invC);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dgam[i].x = fmaf(u[i].x, xh[i].x, dgam[i].x); dgam[i].y = fmaf(u[i].y, xh[i].y, dgam[i].y);
                    dgam[i].z = fmaf(u[i].z, xh[i].z, dgam[i].z); dgam[i].w = fmaf(u[i].w, xh[i].w, dgam[i].w);
                }
                add_row(dbet, u);
                ld_row(lnp, lane, t);
                ln_backward(u, xh, rstd, t, du);
                const int tr = j < d.C ? j : d.C + sid;
                // residual and shared-partial gradients (rows owned by this warp -> no races)
                ld_row(rec + off.R + (size_t)tr * D, lane, t); add_row(t, du); st_row(rec + off.R + (size_t)tr * D, lane, t);
                ld_row(rec + off.G + (size_t)tr * D, lane, t); axpy_row(t, cx.c_w, du); st_row(rec + off.G + (size_t)tr * D, lane, t);
                const float dyy = warp_sum(dot_part(du, ybar));
                ld_row(VFo + (size_t)b * D, lane, t);
                const float d_i = warp_sum(dot_part(du, t));
                ld_row(VFo + (size_t)(d.B + b) * D, lane, t);
                const float d_t = warp_sum(dot_part(du, t));
                ld_row(VFs + (size_t)srow * D, lane, t);
                const float d_s = warp_sum(dot_part(du, t));
                if (lane == 0) {
                    hacc[tr] += cx.c_w * dyy;
                    const float v_i = cx.a_i * (d_i - dyy) * INV_TAU, v_t = cx.a_t * (d_t - dyy) * INV_TAU;
                    dSK[(size_t)b * d.Nsp + cx.r] = v_i;
                    dSK[(size_t)(d.B + b) * d.Nsp + cx.r] = v_t;
                    if (dSKh != nullptr) {
                        dSKh[(size_t)b * d.Nsp + cx.r] = __float2bfloat16_rn(v_i);
                        dSKh[(size_t)(d.B + b) * d.Nsp + cx.r] = __float2bfloat16_rn(v_t);
                    }
                    dtts[tr * 10 + sid] += cx.a_s * (d_s - dyy) * INV_TAU;
                }
                float4 w4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) w4[i] = make_float4(cx.a_i * du[i].x, cx.a_i * du[i].y, cx.a_i * du[i].z, cx.a_i * du[i].w);
                st_row(slot + (0 * TR_WARPS + warp) * D, lane, w4);
#pragma unroll
                for (int i = 0; i < 4; ++i) w4[i] = make_float4(cx.a_t * du[i].x, cx.a_t * du[i].y, cx.a_t * du[i].z, cx.a_t * du[i].w);
                st_row(slot + (1 * TR_WARPS + warp) * D, lane, w4);
#pragma unroll
                for (int i = 0; i < 4; ++i) w4[i] = make_float4(cx.a_s * du[i].x, cx.a_s * du[i].y, cx.a_s * du[i].z, cx.a_s * du[i].w);
                st_row(slot + (2 * TR_WARPS + warp) * D, lane, w4);
            }
            __syncthreads();
            const int nvalid = min(TR_WARPS, d.C + 1 - rd * TR_WARPS);
            for (int w = 0; w < nvalid; ++w) {
                const float2 a = reinterpret_cast<const float2*>(slot + (0 * TR_WARPS + w) * D)[tid];
                const float2 bb = reinterpret_cast<const float2*>(slot + (1 * TR_WARPS + w) * D)[tid];
                const float2 c = reinterpret_cast<const float2*>(slot + (2 * TR_WARPS + w) * D)[tid];
                acc_i.x += a.x; acc_i.y += a.y; acc_t.x += bb.x; acc_t.y += bb.y; acc_s.x += c.x; acc_s.y += c.y;
            }
            __syncthreads();
        }
        reinterpret_cast<float2*>(dVFo + (size_t)b * D)[tid] = acc_i;
        reinterpret_cast<float2*>(dVFo + (size_t)(d.B + b) * D)[tid] = acc_t;
        float2 o = reinterpret_cast<float2*>(dvfst + sid * D)[tid];
        o.x += acc_s.x; o.y += acc_s.y;
        reinterpret_cast<float2*>(dvfst + sid * D)[tid] = o;
    }
    // fold the per-warp LayerNorm gradients in warp order, then publish the record
    __syncthreads();
    st_row(slot + (0 * TR_WARPS + warp) * D, lane, dgam);
    st_row(slot + (1 * TR_WARPS + warp) * D, lane, dbet);
    __syncthreads();
    {
        float2 sg = make_float2(0.f, 0.f), sb = sg;
        for (int w = 0; w < TR_WARPS; ++w) {
            const float2 a = reinterpret_cast<const float2*>(slot + (0 * TR_WARPS + w) * D)[tid];
            const float2 bb = reinterpret_cast<const float2*>(slot + (1 * TR_WARPS + w) * D)[tid];
            sg.x += a.x; sg.y += a.y; sb.x += bb.x; sb.y += bb.y;
        }
        reinterpret_cast<float2*>(rec + off.dgam)[tid] = sg;
        reinterpret_cast<float2*>(rec + off.dbet)[tid] = sb;
    }
    for (int i = tid; i < 10 * D; i += blockDim.x) rec[off.dvfst + i] = dvfst[i];
    for (int i = tid; i < d.Rt; i += blockDim.x) rec[off.h + i] = hacc[i];
    for (int i = tid; i < d.Rt * 10; i += blockDim.x) rec[off.dtts + i] = dtts[i];
}

// out[i] = sum_p partials[p][i] in fixed order (4 interleaved lanes of p, folded in order); up to two
// independent jobs per launch (blocks [0, blocks0) -> job 0, the rest -> job 1).
struct ReduceJob {
    const float* partials;
    float* out;
    size_t len4;
    int n_partials;
};
struct ReduceJobs {
    ReduceJob j[2];
    int blocks0;
};
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const __grid_constant__ ReduceJobs rj) {
    __shared__ float4 fold[4][64];
    const bool second = (int)blockIdx.x >= rj.blocks0;
    const ReduceJob& job = rj.j[second ? 1 : 0];
    const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
    const size_t i = (size_t)((int)blockIdx.x - (second ? rj.blocks0 : 0)) * 64 + c;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < job.len4) {
        for (int p = g; p < job.n_partials; p += 4) {
            const float4 a = reinterpret_cast<const float4*>(job.partials)[(size_t)p * job.len4 + i];
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    fold[g][c] = s;
    __syncthreads();
    if (g == 0 && i < job.len4) {
        float4 r = fold[0][c];
#pragma unroll
        for (int q = 1; q < 4; ++q) { r.x += fold[q][c].x; r.y += fold[q][c].y; r.z += fold[q][c].z; r.w += fold[q][c].w; }
        reinterpret_cast<float4*>(job.out)[i] = r;
    }
}

// compact table-row gradients -> step-row indexed buffers (zeros for prompt / pad rows); block Nsp folds the
// LayerNorm / fc-bias gradients: dgamma/dbeta = table part + own part, dbfc = own part + sum over table rows of R
__global__ void __launch_bounds__(128)
expand_table_kernel(HeadDims d, const float* __restrict__ red, const float* __restrict__ own_red,
                    float* __restrict__ Rfull, float* __restrict__ Gfull, __nv_bfloat16* __restrict__ Gfullh,
                    float* __restrict__ hfull, float* __restrict__ dTT, float* __restrict__ dVFs,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbfc) {
    const TabOff off = tab_offsets(d);
    const int r = blockIdx.x, t = threadIdx.x;
    if (r == d.Nsp) {
        const float4 a = reinterpret_cast<const float4*>(red + off.dgam)[t], b = reinterpret_cast<const float4*>(own_red)[t];
        reinterpret_cast<float4*>(dgamma)[t] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        const float4 c = reinterpret_cast<const float4*>(red + off.dbet)[t], e = reinterpret_cast<const float4*>(own_red + D)[t];
        reinterpret_cast<float4*>(dbeta)[t] = make_float4(c.x + e.x, c.y + e.y, c.z + e.z, c.w + e.w);
        float4 s = reinterpret_cast<const float4*>(own_red + 2 * D)[t];
        for (int tr = 0; tr < d.Rt; ++tr) {
            const float4 q = reinterpret_cast<const float4*>(red + off.R + (size_t)tr * D)[t];
            s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
        }
        reinterpret_cast<float4*>(dbfc)[t] = s;
        return;
    }
    const bool is_state = r >= d.M && r < d.Ns;
    const int tr = r < d.C ? r : (is_state ? d.C + (r - d.M) : -1);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(Rfull + (size_t)r * D)[t] = tr >= 0 ? reinterpret_cast<const float4*>(red + off.R + (size_t)tr * D)[t] : z;
    const float4 gv = tr >= 0 ? reinterpret_cast<const float4*>(red + off.G + (size_t)tr * D)[t] : z;
    reinterpret_cast<float4*>(Gfull + (size_t)r * D)[t] = gv;
    if (Gfullh != nullptr) reinterpret_cast<uint2*>(Gfullh + (size_t)r * D)[t] = pack_bf16x4(gv);
    reinterpret_cast<float4*>(dVFs + (size_t)r * D)[t] = is_state ? reinterpret_cast<const float4*>(red + off.dvfst + (size_t)(r - d.M) * D)[t] : z;
    if (t == 0) hfull[r] = tr >= 0 ? red[off.h + tr] : 0.f;
    for (int j = t; j < d.Nsp; j += blockDim.x)
        dTT[(size_t)r * d.Nsp + j] = (tr >= 0 && j >= d.M && j < d.Ns) ? red[off.dtts + tr * 10 + (j - d.M)] : 0.f;
}

// ------------------------------------------------------------------ own rows: LayerNorm + softmax-output backward
__global__ void __launch_bounds__(256)
ln_own_bwd_kernel(HeadDims d, const float* __restrict__ Ybo, const float* __restrict__ Xo,
                  const float* __restrict__ VFo, const float* __restrict__ aown, const float* __restrict__ bfc,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ g_image,
                  const float* __restrict__ g_text, float* __restrict__ dYo, __nv_bfloat16* __restrict__ dYoh,
                  float* __restrict__ dXo, float* __restrict__ rowdot, float* __restrict__ dsown,
                  float* __restrict__ dVFo, __nv_bfloat16* __restrict__ dVFoh, float* __restrict__ partials) {
    __shared__ __align__(16) float fold[3][8][D];        // 48 KB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 dgam[4], dbet[4], dbf[4];
    zero_row(dgam); zero_row(dbet); zero_row(dbf);
    for (int b = blockIdx.x * 8 + warp; b < d.B; b += gridDim.x * 8) {
        float4 accI[4], accT[4], vi[4], vt[4];
        zero_row(accI); zero_row(accT);
        ld_row(VFo + (size_t)b * D, lane, vi);
        ld_row(VFo + (size_t)(d.B + b) * D, lane, vt);
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const int row = which ? d.B + b : b;
            float4 ybar[4], u[4], t[4], xh[4], du[4];
            ld_row(Ybo + (size_t)row * D, lane, ybar);
            ld_row(Xo + (size_t)row * D, lane, u); add_row(u, ybar);
            ld_row(bfc, lane, t); add_row(u, t);
            const float mean = warp_sum(sum_part(u)) * (1.0f / D);
#pragma unroll
            for (int i = 0; i < 4; ++i) { xh[i].x = u[i].x - mean; xh[i].y = u[i].y - mean; xh[i].z = u[i].z - mean; xh[i].w = u[i].w - mean; }
            const float var = warp_sum(dot_part(xh, xh)) * (1.0f / D);
            const float rstd = 1.0f / sqrtf(var + LN_EPS);
            scale_row(xh, rstd);
            ld_row((which ? g_text : g_image) + (size_t)b * D, lane, u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dgam[i].x = fmaf(u[i].x, xh[i].x, dgam[i].x); dgam[i].y = fmaf(u[i].y, xh[i].y, dgam[i].y);
                dgam[i].z = fmaf(u[i].z, xh[i].z, dgam[i].z); dgam[i].w = fmaf(u[i].w, xh[i].w, dgam[i].w);
            }
            add_row(dbet, u);
            ld_row(gamma, lane, t);
            ln_backward(u, xh, rstd, t, du);
            add_row(dbf, du);
            st_row(dYo + (size_t)row * D, lane, du);
            st_row_h(dYoh != nullptr ? dYoh + (size_t)row * D : nullptr, lane, du);
            st_row(dXo + (size_t)row * D, lane, du);
            const float rd = warp_sum(dot_part(du, ybar));
            const float da_i = warp_sum(dot_part(du, vi));
            const float da_t = warp_sum(dot_part(du, vt));
            const float a0 = aown[2 * row], a1 = aown[2 * row + 1];
            if (lane == 0) {
                rowdot[row] = rd;
                dsown[2 * row] = a0 * (da_i - rd) * INV_TAU;
                dsown[2 * row + 1] = a1 * (da_t - rd) * INV_TAU;
            }
            axpy_row(accI, a0, du);
            axpy_row(accT, a1, du);
        }
        float4 t[4];
        ld_row(dVFo + (size_t)b * D, lane, t); add_row(t, accI); st_row(dVFo + (size_t)b * D, lane, t);
        st_row_h(dVFoh != nullptr ? dVFoh + (size_t)b * D : nullptr, lane, t);
        ld_row(dVFo + (size_t)(d.B + b) * D, lane, t); add_row(t, accT); st_row(dVFo + (size_t)(d.B + b) * D, lane, t);
        st_row_h(dVFoh != nullptr ? dVFoh + (size_t)(d.B + b) * D : nullptr, lane, t);
    }
    st_row(fold[0][warp], lane, dgam);
    st_row(fold[1][warp], lane, dbet);
    st_row(fold[2][warp], lane, dbf);
    __syncthreads();
    float* rec = partials + (size_t)blockIdx.x * OWN_PARTIAL_LEN;
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
        const int q = i / D, c = i % D;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += fold[q][w][c];
        rec[i] = s;
    }
}

// dS = A .* (dA - rowdot) / tau, in place over dA (fp32 + bf16 shadow); 4 columns per thread (Nsp % 16 == 0)
__global__ void __launch_bounds__(256)
ds_kernel(int64_t n4, int Nsp4, const float* __restrict__ Aext, const float* __restrict__ rowdot, float* __restrict__ dA,
          __nv_bfloat16* __restrict__ dSh) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float rd = rowdot[i / Nsp4];
    const float4 a = reinterpret_cast<const float4*>(Aext)[i];
    float4 v = reinterpret_cast<const float4*>(dA)[i];
    v.x = a.x * (v.x - rd) * INV_TAU; v.y = a.y * (v.y - rd) * INV_TAU;
    v.z = a.z * (v.z - rd) * INV_TAU; v.w = a.w * (v.w - rd) * INV_TAU;
    reinterpret_cast<float4*>(dA)[i] = v;
    if (dSh != nullptr) reinterpret_cast<uint2*>(dSh)[i] = pack_bf16x4(v);
}

// own-query x own-key score gradients (the 2x2 per-sample block)
__global__ void __launch_bounds__(256)
own_own_bwd_kernel(HeadDims d, const float* __restrict__ QKVo, const float* __restrict__ dsown,
                   float* __restrict__ dQKVo, __nv_bfloat16* __restrict__ dQKVoh) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= d.B) return;
    const size_t r0 = (size_t)b * 3 * D, r1 = (size_t)(d.B + b) * 3 * D;
    float4 q0[4], q1[4], k0[4], k1[4], t[4];
    ld_row(QKVo + r0, lane, q0); ld_row(QKVo + r1, lane, q1);
    ld_row(QKVo + r0 + D, lane, k0); ld_row(QKVo + r1 + D, lane, k1);
    const float s00 = dsown[2 * b], s01 = dsown[2 * b + 1];
    const float s10 = dsown[2 * (d.B + b)], s11 = dsown[2 * (d.B + b) + 1];
    __nv_bfloat16* const hn = nullptr;
    ld_row(dQKVo + r0, lane, t); axpy_row(t, s00, k0); axpy_row(t, s01, k1); st_row(dQKVo + r0, lane, t);
    st_row_h(dQKVoh != nullptr ? dQKVoh + r0 : hn, lane, t);
    ld_row(dQKVo + r1, lane, t); axpy_row(t, s10, k0); axpy_row(t, s11, k1); st_row(dQKVo + r1, lane, t);
    st_row_h(dQKVoh != nullptr ? dQKVoh + r1 : hn, lane, t);
    ld_row(dQKVo + r0 + D, lane, t); axpy_row(t, s00, q0); axpy_row(t, s10, q1); st_row(dQKVo + r0 + D, lane, t);
    st_row_h(dQKVoh != nullptr ? dQKVoh + r0 + D : hn, lane, t);
    ld_row(dQKVo + r1 + D, lane, t); axpy_row(t, s01, q0); axpy_row(t, s11, q1); st_row(dQKVo + r1 + D, lane, t);
    st_row_h(dQKVoh != nullptr ? dQKVoh + r1 + D : hn, lane, t);
}

// dTT[r][j] += P[r][j] * (GV[r][j] - h[r]) / tau   for j < M; the bf16 shadow is written for every element
__global__ void __launch_bounds__(256)
dtt_kernel(int Nsp, int M, const float* __restrict__ Pt, const float* __restrict__ GV, const float* __restrict__ h,
           float* __restrict__ dTT, __nv_bfloat16* __restrict__ dTTh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nsp * Nsp) return;
    const int r = i / Nsp, j = i % Nsp;
    float v = dTT[i];
    if (j < M) { v += Pt[i] * (GV[i] - h[r]) * INV_TAU; dTT[i] = v; }
    if (dTTh != nullptr) dTTh[i] = __float2bfloat16_rn(v);
}

// dz = inv * (dx - x (x.dx)) for up to four row sets in one launch (fp32 + bf16 shadow), plus per-block column
// sums of dz (fixed order inside the block) for the projection-bias gradients.
//   segment: dz[r] <- from dx[src_row(r)], X[xrow(r)], inv[xrow(r)] with xrow(r) = r + x_off
struct NrmSeg {
    const float* dXsrc;        // [.,D] source gradient rows (row r + x_off)
    const float* X;            // normalised rows
    const float* inv;
    float* dZ;                 // [rows,D] output
    __nv_bfloat16* dZh;
    float* partial;            // [nblocks][D] column sums of this segment's blocks
    int64_t rows;
    int64_t src_off;           // row offset into dXsrc / X / inv
    int rows_per_block;        // multiple of 8
    int blk0;
};
struct NrmList {
    NrmSeg s[4];
    int n;
    int identity;              // 1: dz = dx (no normalisation in the forward)
};
__global__ void __launch_bounds__(256)
nrm_bwd_kernel(const __grid_constant__ NrmList nl) {
    __shared__ __align__(16) float fold[8][D];
    int si = 0;
    for (int q = 1; q < nl.n; ++q)
        if ((int)blockIdx.x >= nl.s[q].blk0) si = q;
    const NrmSeg& sg = nl.s[si];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lb = (int)blockIdx.x - sg.blk0;
    const int64_t r_begin = (int64_t)lb * sg.rows_per_block;
    const int64_t r_end = min(sg.rows, r_begin + sg.rows_per_block);
    float4 acc[4];
    zero_row(acc);
    for (int64_t r = r_begin + warp; r < r_end; r += 8) {
        const int64_t xr = r + sg.src_off;
        float4 x[4], dx[4];
        if (!nl.identity) ld_row(sg.X + xr * D, lane, x);
        ld_row(sg.dXsrc + xr * D, lane, dx);
        if (!nl.identity) {
            const float dt = warp_sum(dot_part(x, dx));
            const float s = sg.inv[xr];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dx[i].x = s * (dx[i].x - x[i].x * dt); dx[i].y = s * (dx[i].y - x[i].y * dt);
                dx[i].z = s * (dx[i].z - x[i].z * dt); dx[i].w = s * (dx[i].w - x[i].w * dt);
            }
        }
        st_row(sg.dZ + r * D, lane, dx);
        st_row_h(sg.dZh != nullptr ? sg.dZh + r * D : nullptr, lane, dx);
        add_row(acc, dx);
    }
    st_row(fold[warp], lane, acc);
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += fold[w][c];
        sg.partial[(size_t)lb * D + c] = s;
    }
}

// last launch of the backward: projection-bias gradients from the nrm_bwd partials (fixed order) and the
// prompt-row gradients (rows [C, C+P) of the step-row gradient).
//   blocks 0..2: b_img = sum(part[0]) + sum(part[2]); b_text = sum(part[1]); b_state = sum(part[3])
//   blocks 3.. : one prompt row each
struct FinishArgs {
    const float* part[4];
    int nblk[4];
    float *b_img, *b_text, *b_state;
    const float* Rfull;
    float* prompts;            // may be null
    int C, P;
};
__global__ void __launch_bounds__(512)
finish_bwd_kernel(const __grid_constant__ FinishArgs fa) {
    __shared__ float4 fold[4][128];
    const int t = threadIdx.x & 127, g = threadIdx.x >> 7;     // 4 groups of 128 threads, each a quarter of the partials
    if (blockIdx.x < 3) {
        const int k = blockIdx.x;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        const int segs[2] = {k == 0 ? 0 : (k == 1 ? 1 : 3), k == 0 ? 2 : -1};
        for (int q = 0; q < 2; ++q) {
            if (segs[q] < 0 || fa.part[segs[q]] == nullptr) continue;
            const float* p = fa.part[segs[q]];
            const int n = fa.nblk[segs[q]];
            const int per = (n + 3) / 4, b0 = g * per, b1 = min(n, b0 + per);
            for (int b = b0; b < b1; ++b) {
                const float4 a = reinterpret_cast<const float4*>(p + (size_t)b * D)[t];
                s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            }
        }
        fold[g][t] = s;
        __syncthreads();
        if (g == 0) {
            float4 r = fold[0][t];
#pragma unroll
            for (int q = 1; q < 4; ++q) { r.x += fold[q][t].x; r.y += fold[q][t].y; r.z += fold[q][t].z; r.w += fold[q][t].w; }
            float* out = k == 0 ? fa.b_img : (k == 1 ? fa.b_text : fa.b_state);
            if (out != nullptr) reinterpret_cast<float4*>(out)[t] = r;
        }
        return;
    }
    // prompt rows: 4 rows per block
    const int r = (blockIdx.x - 3) * 4 + g;
    if (fa.prompts != nullptr && r < fa.P)
        reinterpret_cast<float4*>(fa.prompts + (size_t)r * D)[t] = reinterpret_cast<const float4*>(fa.Rfull + (size_t)(fa.C + r) * D)[t];
}

}  // namespace team
