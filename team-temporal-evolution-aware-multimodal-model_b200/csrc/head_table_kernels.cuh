// Table-query rows (prototype outputs + state output of every sample), second generation: WARP PER SAMPLE.
//
// The first generation (table_rows_fwd_kernel / table_rows_bwd_kernel: one CTA per sample, 4 warps sharing the 21
// rows) was bound by latency, not by work: every row re-read its two table rows (NF_r, S_r) from L2 (L1 hit rate
// 18 % next to 74 KB of shared memory per CTA), 12 warps per SM could not hide it, and the per-sample sums needed a
// cross-warp reduction through shared memory (ncu: IPC 0.84, long-scoreboard 4.3 of 10 stall cycles per issue).
// Here
//   * the Rt = C + 10 table rows (NF_r and S_r + bfc, 4 KB per row) are loaded into shared memory ONCE per CTA
//     (one persistent CTA per SM),
//   * one warp owns one sample: its three value rows stay in registers for all C + 1 query rows and the per-sample
//     sums (sum_r a_i dY, ...) are plain register accumulators - no cross-warp reduction, no block barrier per row,
//   * the batch reductions (LayerNorm gamma/beta gradients, the state-row sums, the scalar coefficient sums) are
//     folded once per round of TW samples by thread-owned columns in a fixed warp order - deterministic, no atomics.
// Needs C + 1 <= 32 (one lane per query row) and Rt <= 32 (shared memory); larger heads use the first generation.
#pragma once
#include "head_bwd_kernels.cuh"

namespace team {

constexpr int TW = 8;               // warps per CTA = samples per round
constexpr int TB_NRS = 12;          // per-row scalar contributions handed to the fold (see table_rows_bwd_kernel: scal)

__host__ __device__ inline bool table2_supported(const HeadDims& d) { return d.C + 1 <= 32 && d.Rt <= 32; }
constexpr int T2_BAR_FLOATS = 4;    // 16 bytes in front of the tables: the mbarrier of the bulk table load
__host__ __device__ inline size_t table2_fwd_smem_floats(const HeadDims& d) { return (size_t)T2_BAR_FLOATS + (size_t)2 * d.Rt * D; }
__host__ __device__ inline size_t table2_bwd_smem_floats(const HeadDims& d) {
    return (size_t)T2_BAR_FLOATS + (size_t)2 * d.Rt * D + (size_t)TW * 3 * D + (size_t)10 * D + (size_t)d.Rt * TQ_NSC + (size_t)TW * 32 * TB_NRS + D + 16;
}

// tabN[tr] = NF_r, tabS[tr] = S_r + bfc;  tr < C: r = tr (prototype row), tr >= C: r = M + tr - C (state-table row).
// The 120 KB come in as THREE bulk copies of the TMA engine (cp.async.bulk, completion on an mbarrier): the prototype rows and
// the state-table rows of NFt, and the s rows Sb = S_r + b_fc that table_prep_kernel writes once per step.  One thread issues
// them, every warp goes on with its sample's own loads and softmax weights and only waits (table2_wait) before the first table
// row is read.  The register-staged load this replaces was a quarter of the kernel at 1 024 samples (ncu stall samples,
// profiles/r2ab_table_rows_bwd2_stall_regions.md) - every CTA reads the same lines at the same moment.
__device__ __forceinline__ uint32_t t2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void table2_load_async(const HeadDims& d, const float* __restrict__ NFt, const float* __restrict__ Sb,
                                                  float* tabN, float* tabS, uint64_t* bar) {
    if (threadIdx.x == 0) {
        const uint32_t b = t2_smem_u32(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t n1 = (uint32_t)d.C * D * 4, n2 = 10u * D * 4, n3 = (uint32_t)d.Rt * D * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n1 + n2 + n3) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(t2_smem_u32(tabN)), "l"(NFt), "r"(n1), "r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(t2_smem_u32(tabN + (size_t)d.C * D)), "l"(NFt + (size_t)d.M * D), "r"(n2), "r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(t2_smem_u32(tabS)), "l"(Sb), "r"(n3), "r"(b) : "memory");
    }
}
// bounded wait for the tables (phase 0 of the barrier; returns at once on every later call)
__device__ __forceinline__ void table2_wait(uint64_t* bar) {
    const uint32_t b = t2_smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(b) : "memory");
        if (ok) return;
        if (spin > (1u << 24)) __trap();
    }
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(TW * 32, 1)
table_rows_fwd2_kernel(HeadDims d, const float* __restrict__ SK, const float* __restrict__ TT,
                       const float* __restrict__ mt, const float* __restrict__ Zt, const float* __restrict__ NFt,
                       const float* __restrict__ VFo, const float* __restrict__ VFs, const float* __restrict__ S,
                       const float* __restrict__ bfc, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const int64_t* __restrict__ state_ids, float* __restrict__ out_proto,
                       float* __restrict__ out_state, const float* __restrict__ Sb) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float t2_smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(t2_smem);
    float* tabN = t2_smem + T2_BAR_FLOATS;
    float* tabS = tabN + (size_t)d.Rt * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    table2_load_async(d, NFt, Sb, tabN, tabS, bar);
    __syncthreads();                                      // the barrier is initialised before anybody polls it
    (void)S; (void)bfc;
    const float invC = 1.0f / (float)d.C;
    for (int b = blockIdx.x * TW + warp; b < d.B; b += gridDim.x * TW) {
        const int sid = clamp_state(state_ids[b]);
        const int srow = d.M + sid;
        float4 vi[4], vt[4], vs[4], acc[4];
        ld_row(VFo + (size_t)b * D, lane, vi);
        ld_row(VFo + (size_t)(d.B + b) * D, lane, vt);
        ld_row(VFs + (size_t)srow * D, lane, vs);
        zero_row(acc);
        TableRowW mine;
        mine.c_w = mine.a_i = mine.a_t = mine.a_s = 0.f; mine.r = 0;
        if (lane <= d.C) mine = table_row_weights(d, b, lane, srow, SK, TT, mt, Zt);
        table2_wait(bar);
        // Prototype rows two at a time: a row is one dependent chain (shared-memory loads -> 512-wide sums -> five shuffle
        // rounds -> rsqrt -> accumulate) and at 1 024 samples a warp owns one sample, so the kernel's time is C + 1 of
        // those chains back to back; two independent rows in flight overlap their shuffle / load latencies.
        int k = 0;
        for (; k + 1 < d.C; k += 2) {
            const TableRowW rwa = shfl_row_weights(mine, k), rwb = shfl_row_weights(mine, k + 1);
            float4 ua[4], ta[4], ub[4], tb[4];
            ld_row(tabN + (size_t)k * D, lane, ua);
            ld_row(tabS + (size_t)k * D, lane, ta);
            ld_row(tabN + (size_t)(k + 1) * D, lane, ub);
            ld_row(tabS + (size_t)(k + 1) * D, lane, tb);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ua[i] = fma4s(rwa.c_w, ua[i], fma4s(rwa.a_i, vi[i], fma4s(rwa.a_t, vt[i], fma4s(rwa.a_s, vs[i], ta[i]))));
                ub[i] = fma4s(rwb.c_w, ub[i], fma4s(rwb.a_i, vi[i], fma4s(rwb.a_t, vt[i], fma4s(rwb.a_s, vs[i], tb[i]))));
            }
            float a1 = sum_part(ua), a2 = dot_part(ua, ua), b1 = sum_part(ub), b2 = dot_part(ub, ub);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                b1 += __shfl_xor_sync(0xffffffffu, b1, o); b2 += __shfl_xor_sync(0xffffffffu, b2, o);
            }
            const float ma = a1 * (1.0f / D), mb = b1 * (1.0f / D);
            const float ra = 1.0f / sqrtf(fmaxf(a2 * (1.0f / D) - ma * ma, 0.f) + LN_EPS);
            const float rb = 1.0f / sqrtf(fmaxf(b2 * (1.0f / D) - mb * mb, 0.f) + LN_EPS);
            shift_row(ua, -ma);
            axpy_row(acc, ra, ua);                 // same order as one row at a time: bit-identical results
            shift_row(ub, -mb);
            axpy_row(acc, rb, ub);
        }
        for (; k <= d.C; ++k) {
            const TableRowW rw = shfl_row_weights(mine, k);
            const int tr = k < d.C ? k : d.C + sid;
            float4 u[4], t[4];
            ld_row(tabN + (size_t)tr * D, lane, u);
            ld_row(tabS + (size_t)tr * D, lane, t);
#pragma unroll
            for (int i = 0; i < 4; ++i) u[i] = fma4s(rw.c_w, u[i], fma4s(rw.a_i, vi[i], fma4s(rw.a_t, vt[i], fma4s(rw.a_s, vs[i], t[i]))));
            // one reduction stage for both LayerNorm statistics (sum and sum of squares, interleaved shuffles):
            // the serial shuffle chains, not the arithmetic, set the time per row of this kernel
            float s1 = sum_part(u), s2 = dot_part(u, u);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            const float mean = s1 * (1.0f / D);
            const float var = fmaxf(s2 * (1.0f / D) - mean * mean, 0.f);
            const float rstd = 1.0f / sqrtf(var + LN_EPS);
            shift_row(u, -mean);
            if (k < d.C) {
                axpy_row(acc, rstd, u);
            } else {
                float4 g[4], be[4];
                ld_row(gamma, lane, g); ld_row(beta, lane, be);
#pragma unroll
                for (int i = 0; i < 4; ++i) u[i] = fma4(mul4s(rstd, u[i]), g[i], be[i]);
                st_row(out_state + (size_t)b * D, lane, u);
            }
        }
        {
            float4 g[4], be[4];
            ld_row(gamma, lane, g); ld_row(beta, lane, be);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fma4(mul4s(invC, acc[i]), g[i], be[i]);
            st_row(out_proto + (size_t)b * D, lane, acc);
        }
    }
}

// ------------------------------------------------------------------ backward
// Same outputs and the same per-CTA partial record as table_rows_bwd_kernel (tab_offsets), see there for the algebra.
// dynamic smem: tabN[Rt][D] | tabS[Rt][D] | slots[TW][3][D] (sum_r a_s dY | dgamma contribution | dbeta contribution)
//               | dvfst[10][D] | scal[Rt][TQ_NSC] | rsl[TW][32][TB_NRS] | gam[D] | sidw[TW]
__global__ void __launch_bounds__(TW * 32, 1)
table_rows_bwd2_kernel(HeadDims d, const float* __restrict__ SK, const float* __restrict__ TT,
                       const float* __restrict__ mt, const float* __restrict__ Zt, const float* __restrict__ NFt,
                       const float* __restrict__ VFo, const float* __restrict__ VFs, const float* __restrict__ S,
                       const float* __restrict__ bfc, const float* __restrict__ gamma,
                       const int64_t* __restrict__ state_ids, const float* __restrict__ g_proto,
                       const float* __restrict__ g_state, float* __restrict__ dSK, __nv_bfloat16* __restrict__ dSKh,
                       float* __restrict__ dVFo, float* __restrict__ GG, __nv_bfloat16* __restrict__ GGh,
                       float* __restrict__ A1, __nv_bfloat16* __restrict__ A1h, float* __restrict__ A23,
                       __nv_bfloat16* __restrict__ A23h, int ldA, float* __restrict__ partials, const float* __restrict__ Sb) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float t2_smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(t2_smem);
    float* tabN = t2_smem + T2_BAR_FLOATS;
    float* tabS = tabN + (size_t)d.Rt * D;
    float* slots = tabS + (size_t)d.Rt * D;                  // [TW][3][D]
    float* dvfst = slots + (size_t)TW * 3 * D;               // [10][D]
    float* scal = dvfst + 10 * D;                            // [Rt][TQ_NSC]
    float* rsl = scal + (size_t)d.Rt * TQ_NSC;               // [TW][32][TB_NRS]
    float* gam = rsl + (size_t)TW * 32 * TB_NRS;             // [D]
    int* sidw = reinterpret_cast<int*>(gam + D);             // [TW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TabOff off = tab_offsets(d);
    const int gcol = ldA / 2;
    table2_load_async(d, NFt, Sb, tabN, tabS, bar);
    (void)S; (void)bfc;
    for (int i = tid; i < 10 * D; i += blockDim.x) dvfst[i] = 0.f;
    for (int i = tid; i < d.Rt * TQ_NSC; i += blockDim.x) scal[i] = 0.f;
    for (int i = tid; i < D; i += blockDim.x) gam[i] = gamma[i];
    float2 dgam = make_float2(0.f, 0.f), dbet = dgam;        // float2 column tid of the 512-wide rows
    const float invC = d.C > 1 ? 1.0f / (float)d.C : 1.0f;
    const __nv_bfloat16 hz = __float2bfloat16_rn(0.f);
    __syncthreads();
    const int ngroups = (d.B + TW - 1) / TW;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int b = grp * TW + warp;
        if (b < d.B) {
            const int sid = clamp_state(state_ids[b]);
            const int srow = d.M + sid;
            float4 vi[4], vt[4], vs[4], ggp[4];
            ld_row(VFo + (size_t)b * D, lane, vi);
            ld_row(VFo + (size_t)(d.B + b) * D, lane, vt);
            ld_row(VFs + (size_t)srow * D, lane, vs);
            float m1p, m1s;
            {   // GG rows (cotangent .* gamma) for the coefficient GEMM; zero the sparse rows of this sample
                float4 g4[4], gs[4];
                ld_row(gam, lane, g4);
                if (g_proto != nullptr) ld_row(g_proto + (size_t)b * D, lane, ggp); else zero_row(ggp);
                ld_row(g_state + (size_t)b * D, lane, gs);
#pragma unroll
                for (int i = 0; i < 4; ++i) { ggp[i] = mul4(mul4s(invC, ggp[i]), g4[i]); gs[i] = mul4(gs[i], g4[i]); }
                st_row(GG + (size_t)b * D, lane, ggp);
                st_row(GG + (size_t)(d.B + b) * D, lane, gs);
                if (GGh != nullptr) { st_row_h(GGh + (size_t)b * D, lane, ggp); st_row_h(GGh + (size_t)(d.B + b) * D, lane, gs); }
                m1p = warp_sum(sum_part(ggp)) * (1.0f / D);
                m1s = warp_sum(sum_part(gs)) * (1.0f / D);
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
            }
            __syncwarp();                                     // the zero fill is ordered before the per-row entries below
            TableRowW mine;
            mine.c_w = mine.a_i = mine.a_t = mine.a_s = 0.f; mine.r = 0;
            if (lane <= d.C) mine = table_row_weights(d, b, lane, srow, SK, TT, mt, Zt);
            table2_wait(bar);
            float4 acc_i[4], acc_t[4], acc_s[4], xs[4];
            zero_row(acc_i); zero_row(acc_t); zero_row(acc_s); zero_row(xs);
            float k_alpha = 0.f, k_beta = 0.f, k_mean = 0.f, k_m1 = 0.f, k_yy = 0.f, k_i = 0.f, k_t = 0.f, k_s = 0.f;
            // g_proto == NULL (the learner's losses never touch the prototype output, models/proof.py:434-442): the C
            // prototype rows have a zero cotangent, so only the state row is differentiated
            for (int k = g_proto != nullptr ? 0 : d.C; k <= d.C; ++k) {
                const TableRowW rw = shfl_row_weights(mine, k);
                const bool is_proto = k < d.C;
                const int tr = is_proto ? k : d.C + sid;
                float4 ybar[4], xh[4], gg[4];
                ld_row(tabN + (size_t)tr * D, lane, ybar);
#pragma unroll
                for (int i = 0; i < 4; ++i) ybar[i] = fma4s(rw.c_w, ybar[i], fma4s(rw.a_i, vi[i], fma4s(rw.a_t, vt[i], mul4s(rw.a_s, vs[i]))));
                ld_row(tabS + (size_t)tr * D, lane, xh);
                add_row(xh, ybar);
                float m1;
                if (is_proto) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) gg[i] = ggp[i];
                    m1 = m1p;
                } else {
                    float4 g4[4];
                    ld_row(gam, lane, g4);
                    ld_row(g_state + (size_t)b * D, lane, gg);
#pragma unroll
                    for (int i = 0; i < 4; ++i) gg[i] = mul4(gg[i], g4[i]);
                    m1 = m1s;
                }
                // ONE reduction stage for mean, variance and m2 = mean(gg .* xhat): sum u, sum u^2, sum gg u (interleaved
                // shuffles); m2 = rstd (sum gg u - mean sum gg) / D with sum gg = D m1.  The serial shuffle chains set the
                // time per row here, so three stages folded into one is worth more than the extra multiply-adds.
                float s1 = sum_part(xh), s2 = dot_part(xh, xh), s3 = dot_part(gg, xh);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
                }
                const float mean = s1 * (1.0f / D);
                const float var = fmaxf(s2 * (1.0f / D) - mean * mean, 0.f);
                const float rstd = 1.0f / sqrtf(var + LN_EPS);
                const float m2 = rstd * (s3 * (1.0f / D) - mean * m1);
#pragma unroll
                for (int i = 0; i < 4; ++i) xh[i] = mul4s(rstd, add4s(-mean, xh[i]));
#pragma unroll
                for (int i = 0; i < 4; ++i) gg[i] = mul4s(rstd, fma4s(-m2, xh[i], add4s(-m1, gg[i])));     // dY
                float p_yy = dot_part(gg, ybar), p_i = dot_part(gg, vi), p_t = dot_part(gg, vt), p_s = dot_part(gg, vs);
                axpy_row(acc_i, rw.a_i, gg);
                axpy_row(acc_t, rw.a_t, gg);
                axpy_row(acc_s, rw.a_s, gg);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {            // four reductions interleaved
                    p_yy += __shfl_xor_sync(0xffffffffu, p_yy, o); p_i += __shfl_xor_sync(0xffffffffu, p_i, o);
                    p_t += __shfl_xor_sync(0xffffffffu, p_t, o); p_s += __shfl_xor_sync(0xffffffffu, p_s, o);
                }
                if (is_proto) {
                    add_row(xs, xh);
                } else {
                    // last row: dgamma / dbeta contributions of this sample: gp .* sum_{j<C} xhat_j + gs .* xhat_state, gp C + gs
                    float4 gp[4], gs[4];
                    if (g_proto != nullptr) ld_row(g_proto + (size_t)b * D, lane, gp); else zero_row(gp);
                    ld_row(g_state + (size_t)b * D, lane, gs);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        gp[i] = mul4s(invC, gp[i]);
                        xs[i] = fma4(gp[i], xs[i], mul4(gs[i], xh[i]));
                        gs[i] = fma4s((float)d.C, gp[i], gs[i]);
                    }
                    st_row(slots + ((size_t)warp * 3 + 1) * D, lane, xs);
                    st_row(slots + ((size_t)warp * 3 + 2) * D, lane, gs);
                }
                if (lane == k) {                              // lane k keeps the (warp-uniform) scalars of its row
                    k_alpha = rstd; k_beta = rstd * rstd * m2; k_mean = mean; k_m1 = m1;
                    k_yy = p_yy; k_i = p_i; k_t = p_t; k_s = p_s;
                }
            }
            st_row(dVFo + (size_t)b * D, lane, acc_i);
            st_row(dVFo + (size_t)(d.B + b) * D, lane, acc_t);
            st_row(slots + (size_t)warp * 3 * D, lane, acc_s);
            if (lane <= d.C) {                               // scalar outputs, one row per lane
                const bool is_proto = lane < d.C;
                const int tr = is_proto ? lane : d.C + sid;
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
        for (int w = 0; w < TW; ++w) {
            const int s = sidw[w];
            if (s < 0) continue;
            const float2 as = reinterpret_cast<const float2*>(slots + ((size_t)w * 3 + 0) * D)[tid];
            const float2 cg = reinterpret_cast<const float2*>(slots + ((size_t)w * 3 + 1) * D)[tid];
            const float2 cb = reinterpret_cast<const float2*>(slots + ((size_t)w * 3 + 2) * D)[tid];
            dgam.x += cg.x; dgam.y += cg.y; dbet.x += cb.x; dbet.y += cb.y;
            float2* dv = reinterpret_cast<float2*>(dvfst + (size_t)s * D) + tid;
            float2 o = *dv;
            o.x += as.x; o.y += as.y;
            *dv = o;
        }
        for (int t = tid; t < (d.C + 1) * TB_NRS; t += blockDim.x) {
            const int j = t / TB_NRS, kk = t - j * TB_NRS;
            for (int w = 0; w < TW; ++w) {
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
    reinterpret_cast<float2*>(rec + off.dgam)[tid] = dgam;
    reinterpret_cast<float2*>(rec + off.dbet)[tid] = dbet;
    for (int i = tid; i < d.Rt * TQ_NSC; i += blockDim.x) rec[off.scal + i] = scal[i];
    for (size_t i = off.scal + (size_t)d.Rt * TQ_NSC + tid; i < off.len; i += blockDim.x) rec[i] = 0.f;
}

}  // namespace team
