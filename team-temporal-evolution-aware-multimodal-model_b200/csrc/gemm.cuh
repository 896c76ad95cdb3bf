// Internal GEMM entry points shared by the head orchestration.
#pragma once
#include "common.cuh"

namespace team {

// fp32 SIMT GEMM: C = alpha*op(A)op(B) + beta*C (+bias[N]).  ta: A stored [K,M]; tb: B stored [N,K].
int gemm_f32(cudaStream_t st, bool ta, bool tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias,
             void* ws, size_t ws_bytes);
size_t gemm_f32_workspace_bytes(int64_t M, int64_t N, int64_t K);

}  // namespace team
