// fp32 FFMA GEMM for the parity mode (TEAM_MODE_F32, <= 1e-5) and for the small per-step
// GEMMs on the <= ~140 shared rows.  C[M,N] = alpha * op(A) op(B) + beta * C (+ bias[N]).
// 64x64x16 tiles, 256 threads, 4x4 micro-tile, 128-bit global loads along the contiguous
// dimension, optional split-K with a fixed-order (deterministic) second pass.
#include "common.cuh"
#include "gemm.cuh"
#include "prof.cuh"

namespace team {

constexpr int GBM = 64, GBN = 64, GBK = 16, GPAD = 4;

// Load 4 consecutive floats starting at p (element index `i` of a run of length `n` valid).
__device__ __forceinline__ float4 ld4_guard(const float* p, int64_t i, int64_t n, bool row_ok, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!row_ok) return v;
    if (vec_ok && i + 3 < n) return *reinterpret_cast<const float4*>(p);
    if (i < n) v.x = p[0];
    if (i + 1 < n) v.y = p[1];
    if (i + 2 < n) v.z = p[2];
    if (i + 3 < n) v.w = p[3];
    return v;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda,
                const float* __restrict__ B, int64_t ldb, float beta, float* __restrict__ C, int64_t ldc,
                const float* __restrict__ bias, int k_per_split, float* __restrict__ partial) {
    __shared__ __align__(16) float As[GBK][GBM + GPAD];
    __shared__ __align__(16) float Bs[GBK][GBN + GPAD];
    const int t = threadIdx.x;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int tx = t & 15, ty = t >> 4;
    const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += GBK) {
        // ---- A tile -> As[k][m]
        if (!TA) {      // A[m][k], k contiguous: thread = (row t/4, k-quad t%4)
            const int row = t >> 2, kq = t & 3;
            const int m = m0 + row, k = k0 + 4 * kq;
            const float4 v = ld4_guard(A + (int64_t)m * lda + k, k, kend, m < M, a_vec);
            As[4 * kq + 0][row] = v.x; As[4 * kq + 1][row] = v.y; As[4 * kq + 2][row] = v.z; As[4 * kq + 3][row] = v.w;
        } else {        // A[k][m], m contiguous: thread = (k t/16, m-quad t%16)
            const int kk = t >> 4, mq = t & 15;
            const int k = k0 + kk, m = m0 + 4 * mq;
            const float4 v = ld4_guard(A + (int64_t)k * lda + m, m, M, k < kend, a_vec);
            *reinterpret_cast<float4*>(&As[kk][4 * mq]) = v;
        }
        // ---- B tile -> Bs[k][n]
        if (TB) {       // B[n][k], k contiguous
            const int row = t >> 2, kq = t & 3;
            const int n = n0 + row, k = k0 + 4 * kq;
            const float4 v = ld4_guard(B + (int64_t)n * ldb + k, k, kend, n < N, b_vec);
            Bs[4 * kq + 0][row] = v.x; Bs[4 * kq + 1][row] = v.y; Bs[4 * kq + 2][row] = v.z; Bs[4 * kq + 3][row] = v.w;
        } else {        // B[k][n], n contiguous
            const int kk = t >> 4, nq = t & 15;
            const int k = k0 + kk, n = n0 + 4 * nq;
            const float4 v = ld4_guard(B + (int64_t)k * ldb + n, n, N, k < kend, b_vec);
            *reinterpret_cast<float4*>(&Bs[kk][4 * nq]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (partial != nullptr) {       // split-K: raw partial sums, compact [z][M][N]
        float* P = partial + (size_t)blockIdx.z * M * N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + 4 * ty + i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + 4 * tx + j;
                if (n < N) P[(size_t)m * N + n] = acc[i][j];
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * ty + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + 4 * tx + j;
            if (n >= N) continue;
            float v = alpha * acc[i][j];
            if (bias != nullptr) v += bias[n];
            if (beta != 0.f) v += beta * C[(int64_t)m * ldc + n];
            C[(int64_t)m * ldc + n] = v;
        }
    }
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, float alpha, float beta,
                     float* __restrict__ C, int64_t ldc, const float* __restrict__ bias) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    const int m = (int)(idx / N), n = (int)(idx % N);
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * M * N + idx];      // fixed order
    float v = alpha * s;
    if (bias != nullptr) v += bias[n];
    if (beta != 0.f) v += beta * C[(int64_t)m * ldc + n];
    C[(int64_t)m * ldc + n] = v;
}

size_t gemm_f32_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ((M + GBM - 1) / GBM) * ((N + GBN - 1) / GBN);
    if (tiles >= NUM_SMS || K < 512) return 0;
    int64_t splits = (2 * NUM_SMS + tiles - 1) / tiles;
    if (splits > K / 128) splits = K / 128;
    if (splits < 2) return 0;
    return (size_t)splits * M * N * sizeof(float);
}

int gemm_f32(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias,
             void* ws, size_t ws_bytes) {
    if (M <= 0 || N <= 0) return TEAM_OK;
    TEAM_REQUIRE(K >= 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm_f32: bad shape");
    const int64_t tiles = ((M + GBM - 1) / GBM) * ((N + GBN - 1) / GBN);
    int splits = 1;
    if (tiles < NUM_SMS && K >= 512 && ws != nullptr) {
        int64_t s = (2 * NUM_SMS + tiles - 1) / tiles;
        if (s > K / 128) s = K / 128;
        if (s >= 2 && (size_t)s * M * N * sizeof(float) <= ws_bytes) splits = (int)s;
    }
    int kps = (int)((K + splits - 1) / splits);
    kps = (kps + GBK - 1) / GBK * GBK;
    splits = (int)((K + kps - 1) / kps);
    if (splits < 1) splits = 1;
    float* partial = splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
    dim3 grid((unsigned)((N + GBN - 1) / GBN), (unsigned)((M + GBM - 1) / GBM), (unsigned)splits);
#define TEAM_GEMM_LAUNCH(TA_, TB_)                                                                       \
    gemm_f32_kernel<TA_, TB_><<<grid, 256, 0, st>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, \
                                                    C, ldc, bias, kps, partial)
    const int pslot = prof_enabled() ? prof_begin(st, 0, 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)) : -1;
    if (!ta && tb) TEAM_GEMM_LAUNCH(false, true);
    else if (!ta && !tb) TEAM_GEMM_LAUNCH(false, false);
    else if (ta && !tb) TEAM_GEMM_LAUNCH(true, false);
    else TEAM_GEMM_LAUNCH(true, true);
#undef TEAM_GEMM_LAUNCH
    if (pslot >= 0) prof_end(st, pslot);
    TEAM_LAUNCH_CHECK("gemm_f32_kernel");
    if (splits > 1) {
        const int64_t tot = M * N;
        splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(partial, splits, (int)M, (int)N, alpha, beta, C, ldc, bias);
        TEAM_LAUNCH_CHECK("splitk_reduce_kernel");
    }
    return TEAM_OK;
}

}  // namespace team

extern "C" int team_gemm_f32(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                             const float* bias, void* workspace, size_t workspace_bytes, void* stream) {
    return team::gemm_f32((cudaStream_t)stream, ta != 0, tb != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc,
                          bias, workspace, workspace_bytes);
}
