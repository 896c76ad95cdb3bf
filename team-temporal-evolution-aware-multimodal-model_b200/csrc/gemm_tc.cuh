// tcgen05 / TMA / TMEM bf16 GEMM (mode BF16) - declarations.
#pragma once
#include "head.cuh"

namespace team {

// bytes of bf16 operand staging the BF16 mode needs inside the head workspace
size_t tc_operand_bytes(const HeadDims& d);

}  // namespace team
