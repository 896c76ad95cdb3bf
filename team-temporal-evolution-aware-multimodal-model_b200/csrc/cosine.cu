// K8 - fused L2-normalise + cosine logits + argmax against <= 32 class rows per pass.
//
// Replaces CosineLinear.forward (convs/linears.py:51-61) and the last stage of
// Learner.forward_for_classification (models/proof.py:526-535).
//
// HBM-bound for the class counts of the path (C <= 20): each feature row (2 KB fp32) is
// read once with 128-bit streaming loads; the normalised class matrix lives in shared
// memory.  A warp owns two rows at a time: lane l keeps columns {4(l+32j)..+3} of both rows
// in registers, accumulates one partial dot product per class, and a 31-shuffle
// transpose-reduce leaves the total for class l on lane l (instead of 5 shuffles per
// class).  Algorithmic bytes per row: 512*e + 4*C (logits) + 8 (argmax)   (SURVEY 8d).
#include "common.cuh"
#include "head.cuh"

namespace team {

constexpr int COS_WARPS = 8;
constexpr int COS_CCHUNK = 32;

template <typename T> struct XLoad;
template <> struct XLoad<float> {
    static __device__ __forceinline__ float4 load(const float* row, int c4) { return ld_stream_f4(row + 4 * c4); }
};
template <> struct XLoad<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const __nv_bfloat16* row, int c4) {
        uint2 u = ld_stream_u2(row + 4 * c4);
        return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    }
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

// v[0..31] partial sums per class on every lane -> returns the full sum of class `lane`.
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <typename T>
__global__ void __launch_bounds__(COS_WARPS * 32)
cosine_logits_kernel(const T* __restrict__ x, int64_t n_rows, const float* __restrict__ w, int num_classes,
                     const float* __restrict__ sigma_dev, float* __restrict__ logits,
                     int64_t* __restrict__ argmax_out, float* __restrict__ chunk_best, int n_chunks) {
    extern __shared__ __align__(16) float4 ws[];         // [min(C,32)][128] normalised class rows of this chunk
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.y * COS_CCHUNK;
    const int nc = min(COS_CCHUNK, num_classes - c0);
    const int c_alloc = min(COS_CCHUNK, num_classes);
    for (int c = warp; c < c_alloc; c += COS_WARPS) {
        float4 r[4];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            r[j] = (c < nc) ? reinterpret_cast<const float4*>(w + (size_t)(c0 + c) * D)[lane + 32 * j]
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += dot4(r[j], r[j]);
        }
        ss = warp_sum(ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), NORM_EPS);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            ws[c * (D / 4) + lane + 32 * j] = make_float4(r[j].x * inv, r[j].y * inv, r[j].z * inv, r[j].w * inv);
    }
    __syncthreads();
    const float sigma = sigma_dev ? __ldg(sigma_dev) : 1.0f;
    const int64_t pair_stride = (int64_t)gridDim.x * COS_WARPS;
    const int64_t n_pairs = (n_rows + 1) / 2;
    for (int64_t pr = (int64_t)blockIdx.x * COS_WARPS + warp; pr < n_pairs; pr += pair_stride) {
        const int64_t ra = 2 * pr, rb = 2 * pr + 1;
        const bool has_b = rb < n_rows;
        float4 xa[4], xb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            xa[j] = XLoad<T>::load(x + ra * D, lane + 32 * j);
            xb[j] = has_b ? XLoad<T>::load(x + rb * D, lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float ssa = 0.f, ssb = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { ssa += dot4(xa[j], xa[j]); ssb += dot4(xb[j], xb[j]); }
        ssa = warp_sum(ssa);
        ssb = warp_sum(ssb);
        float va[32], vb[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float a = 0.f, b = 0.f;
            if (c < nc) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w4 = ws[c * (D / 4) + lane + 32 * j];
                    a += dot4(xa[j], w4);
                    b += dot4(xb[j], w4);
                }
            }
            va[c] = a;
            vb[c] = b;
        }
        float ta = transpose_reduce32(va, lane);
        float tb = transpose_reduce32(vb, lane);
        ta *= sigma / fmaxf(sqrtf(ssa), NORM_EPS);
        tb *= sigma / fmaxf(sqrtf(ssb), NORM_EPS);
        if (logits != nullptr && lane < nc) {
            logits[ra * num_classes + c0 + lane] = ta;
            if (has_b) logits[rb * num_classes + c0 + lane] = tb;
        }
        if (argmax_out != nullptr) {
            // first maximal index (torch.max semantics): max value, ties -> lower index
            float ba = lane < nc ? ta : -INFINITY, bb = lane < nc ? tb : -INFINITY;
            int ia = lane, ib = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float oa = __shfl_xor_sync(0xffffffffu, ba, o);
                const int oia = __shfl_xor_sync(0xffffffffu, ia, o);
                if (oa > ba || (oa == ba && oia < ia)) { ba = oa; ia = oia; }
                const float ob = __shfl_xor_sync(0xffffffffu, bb, o);
                const int oib = __shfl_xor_sync(0xffffffffu, ib, o);
                if (ob > bb || (ob == bb && oib < ib)) { bb = ob; ib = oib; }
            }
            if (lane == 0) {
                if (n_chunks == 1) {
                    argmax_out[ra] = c0 + ia;
                    if (has_b) argmax_out[rb] = c0 + ib;
                } else {   // multi-chunk: stash (value,index) per chunk, merged by cosine_argmax_merge
                    chunk_best[(ra * n_chunks + blockIdx.y) * 2 + 0] = ba;
                    chunk_best[(ra * n_chunks + blockIdx.y) * 2 + 1] = __int_as_float(c0 + ia);
                    if (has_b) {
                        chunk_best[(rb * n_chunks + blockIdx.y) * 2 + 0] = bb;
                        chunk_best[(rb * n_chunks + blockIdx.y) * 2 + 1] = __int_as_float(c0 + ib);
                    }
                }
            }
        }
    }
}

__global__ void cosine_argmax_merge(const float* __restrict__ chunk_best, int64_t n_rows, int n_chunks,
                                    int64_t* __restrict__ argmax_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    float best = -INFINITY;
    int bi = 0;
    for (int q = 0; q < n_chunks; ++q) {
        const float v = chunk_best[(r * n_chunks + q) * 2];
        const int i = __float_as_int(chunk_best[(r * n_chunks + q) * 2 + 1]);
        if (q == 0 || v > best) { best = v; bi = i; }
    }
    argmax_out[r] = bi;
}

}  // namespace team

using namespace team;

int team::cosine_logits_launch(cudaStream_t st, const float* x, int64_t n_rows, const float* w, int64_t num_classes,
                               const float* sigma_dev, float* logits, int64_t* argmax) {
    return team_cosine_logits(x, TEAM_DTYPE_F32, n_rows, w, num_classes, sigma_dev, logits, argmax, (void*)st);
}

extern "C" int team_cosine_logits(const void* x, int x_dtype, int64_t n_rows, const float* w,
                                  int64_t num_classes, const float* sigma_dev, float* logits,
                                  int64_t* argmax, void* stream) {
    TEAM_REQUIRE(x != nullptr && w != nullptr && n_rows >= 0 && num_classes >= 1, "team_cosine_logits: bad args");
    TEAM_REQUIRE(x_dtype == TEAM_DTYPE_F32 || x_dtype == TEAM_DTYPE_BF16, "team_cosine_logits: bad dtype %d", x_dtype);
    if (n_rows == 0) return TEAM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_chunks = (int)((num_classes + COS_CCHUNK - 1) / COS_CCHUNK);
    float* chunk_best = nullptr;
    if (n_chunks > 1 && argmax != nullptr) {
        // rare path (C > 32): scratch from the stream-ordered allocator
        TEAM_CUDA_CHECK(cudaMallocAsync((void**)&chunk_best, (size_t)n_rows * n_chunks * 2 * sizeof(float), st));
    }
    const int64_t n_pairs = (n_rows + 1) / 2;
    int64_t gx = (n_pairs + COS_WARPS - 1) / COS_WARPS;
    const int64_t cap = (int64_t)NUM_SMS * 2;          // persistent over row pairs (register-limited to 2 CTAs / SM)
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, (unsigned)n_chunks, 1);
    const size_t smem = (size_t)(num_classes < COS_CCHUNK ? num_classes : COS_CCHUNK) * D * sizeof(float);
    if (x_dtype == TEAM_DTYPE_F32) {
        TEAM_CUDA_CHECK(cudaFuncSetAttribute(cosine_logits_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cosine_logits_kernel<float><<<grid, COS_WARPS * 32, smem, st>>>(reinterpret_cast<const float*>(x), n_rows, w, (int)num_classes, sigma_dev, logits, argmax, chunk_best, n_chunks);
    } else {
        TEAM_CUDA_CHECK(cudaFuncSetAttribute(cosine_logits_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cosine_logits_kernel<__nv_bfloat16><<<grid, COS_WARPS * 32, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), n_rows, w, (int)num_classes, sigma_dev, logits, argmax, chunk_best, n_chunks);
    }
    TEAM_LAUNCH_CHECK("cosine_logits_kernel");
    if (chunk_best != nullptr) {
        cosine_argmax_merge<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(chunk_best, n_rows, n_chunks, argmax);
        TEAM_LAUNCH_CHECK("cosine_argmax_merge");
        TEAM_CUDA_CHECK(cudaFreeAsync(chunk_best, st));
    }
    return TEAM_OK;
}
