// Per-row / per-sample CUDA-core kernels of the fusion head (softmax, LayerNorm, table-query
// rows, normalisation) - warp-per-row with 128-bit loads and shuffle reductions.
// Lane l of a warp owns float4 columns {l, l+32, l+64, l+96} of a 512-wide row.
#pragma once
#include "head.cuh"

namespace team {

constexpr float INV_TAU = 0.04419417382415922f;      // 1/sqrt(512)  (convs/projections.py:57)
constexpr int TR_WARPS = 8;                           // warps per CTA in the table-row kernels (thread t owns columns 2t,2t+1)

struct PtrList {
    const float* p[TEAM_MAX_TASKS];
    int n;
};

__device__ __forceinline__ void ld_row(const float* p, int lane, float4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = reinterpret_cast<const float4*>(p)[lane + 32 * i];
}
__device__ __forceinline__ void st_row(float* p, int lane, const float4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[lane + 32 * i] = v[i];
}
// bf16 shadow of a 512-wide row (same lane -> column mapping); p may be null (F32 mode)
__device__ __forceinline__ void st_row_h(__nv_bfloat16* p, int lane, const float4 (&v)[4]) {
    if (p == nullptr) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<uint2*>(p)[lane + 32 * i] = pack_bf16x4(v[i]);
}
__device__ __forceinline__ float dot_part(const float4 (&a)[4], const float4 (&b)[4]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += a[i].x * b[i].x + a[i].y * b[i].y + a[i].z * b[i].z + a[i].w * b[i].w;
    return s;
}
__device__ __forceinline__ float sum_part(const float4 (&a)[4]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += a[i].x + a[i].y + a[i].z + a[i].w;
    return s;
}
__device__ __forceinline__ void axpy_row(float4 (&y)[4], float a, const float4 (&x)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        y[i].x = fmaf(a, x[i].x, y[i].x); y[i].y = fmaf(a, x[i].y, y[i].y);
        y[i].z = fmaf(a, x[i].z, y[i].z); y[i].w = fmaf(a, x[i].w, y[i].w);
    }
}
__device__ __forceinline__ void add_row(float4 (&y)[4], const float4 (&x)[4]) { axpy_row(y, 1.0f, x); }
__device__ __forceinline__ void scale_row(float4 (&y)[4], float a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { y[i].x *= a; y[i].y *= a; y[i].z *= a; y[i].w *= a; }
}
__device__ __forceinline__ void zero_row(float4 (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// LayerNorm over a 512-wide row held by one warp (eps 1e-5, biased variance).
__device__ __forceinline__ void ln_forward(const float4 (&u)[4], const float4 (&g)[4], const float4 (&be)[4],
                                           float4 (&xh)[4], float& rstd, float4 (&o)[4]) {
    const float mean = warp_sum(sum_part(u)) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        xh[i].x = u[i].x - mean; xh[i].y = u[i].y - mean; xh[i].z = u[i].z - mean; xh[i].w = u[i].w - mean;
    }
    const float var = warp_sum(dot_part(xh, xh)) * (1.0f / D);
    rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        xh[i].x *= rstd; xh[i].y *= rstd; xh[i].z *= rstd; xh[i].w *= rstd;
        o[i].x = fmaf(xh[i].x, g[i].x, be[i].x); o[i].y = fmaf(xh[i].y, g[i].y, be[i].y);
        o[i].z = fmaf(xh[i].z, g[i].z, be[i].z); o[i].w = fmaf(xh[i].w, g[i].w, be[i].w);
    }
}
// du = rstd * (gg - mean(gg) - xh * mean(gg*xh)),  gg = go * gamma
__device__ __forceinline__ void ln_backward(const float4 (&go)[4], const float4 (&xh)[4], float rstd,
                                            const float4 (&g)[4], float4 (&du)[4]) {
    float4 gg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gg[i] = make_float4(go[i].x * g[i].x, go[i].y * g[i].y, go[i].z * g[i].z, go[i].w * g[i].w);
    const float m1 = warp_sum(sum_part(gg)) * (1.0f / D);
    const float m2 = warp_sum(dot_part(gg, xh)) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        du[i].x = rstd * (gg[i].x - m1 - xh[i].x * m2); du[i].y = rstd * (gg[i].y - m1 - xh[i].y * m2);
        du[i].z = rstd * (gg[i].z - m1 - xh[i].z * m2); du[i].w = rstd * (gg[i].w - m1 - xh[i].w * m2);
    }
}

__device__ __forceinline__ int clamp_state(int64_t s) { return s < 0 ? 0 : (s > 9 ? 9 : (int)s); }

}  // namespace team
