// Optional per-launch CUDA-event timing of the GEMM kernels (bench.py's roofline numbers).
// Disabled by default; never used under stream capture.
#pragma once
#include "common.cuh"

namespace team {
bool prof_enabled();
// record an event pair around one launch on `st`; returns slot or -1
int prof_begin(cudaStream_t st, int kind, double flops, double bytes);
void prof_end(cudaStream_t st, int slot);
}  // namespace team
