// tcgen05 / TMA / TMEM bf16 GEMM (mode BF16) - declarations.
#pragma once
#include "head.cuh"

namespace team {

struct TcGemm {
    bool a_mn, b_mn;          // false: operand stored [rows,K] (K-major); true: stored [K,rows] (MN-major)
    int64_t M, N, K;
    float alpha, beta;
    const void* A;            // bf16
    const void* A2;           // optional bf16 "lo" half of A (same layout), may be null
    int64_t lda;
    const void* B;            // bf16
    int64_t ldb;
    float* C;                 // fp32 [M,N]
    int64_t ldc;
    const float* bias;        // fp32 [N] or null
};

int gemm_bf16_tc(cudaStream_t st, const TcGemm& g, void* ws, size_t ws_bytes);
// fp32 [rows,cols] (lds) -> bf16 hi (+ optional lo = bf16(x - hi)) with leading dimension ldd
int to_bf16(cudaStream_t st, const float* src, int64_t lds, int64_t rows, int cols, void* hi, void* lo, int64_t ldd);
// bytes of bf16 operand staging the BF16 mode needs inside the head workspace
size_t tc_operand_bytes(const HeadDims& d);

}  // namespace team
