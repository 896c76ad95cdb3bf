#include "gemm_tc.cuh"
namespace team {
size_t tc_operand_bytes(const HeadDims& d) { (void)d; return 0; }
}
extern "C" int team_gemm_bf16_nt(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                                 int64_t ldb, float* C, int64_t ldc, void* stream) {
    team::set_error("team_gemm_bf16_nt: not built yet");
    return TEAM_EUNSUPPORTED;
}
