// tcgen05 / TMA / TMEM bf16 GEMM for sm_100a (TEAM_MODE_BF16).
//
//   C[M,N] (fp32) = alpha * op(A) op(B) (+ bias[N]) (+ beta * C)      fp32 accumulation in TMEM
//
// Operands are bf16 in global memory and reach shared memory through TMA
// (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a multi-stage mbarrier ring; ONE elected
// thread issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1); the 128 x BN fp32
// accumulator lives in tensor memory and is read back with tcgen05.ld by the 4 epilogue
// warps (warp w owns TMEM lanes 32w..32w+31 = output rows).  Either operand may be K-major
// (row-major [rows,K]) or MN-major (row-major [K,rows]) - selected by the UMMA instruction
// descriptor major bits and the matching shared-memory descriptor - so the backward's
// A^T B (weight gradients, reductions over the batch) and A B products need no transposed
// copies.  Split-K (grid.z) writes fp32 partials that are folded in a fixed order.
// An optional second A tensor ("lo" half of a 2-term bf16 split of an fp32 activation) is
// accumulated into the same tile: C = (A_hi + A_lo) B, keeping 16 mantissa bits.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include "gemm_tc.cuh"
#include "prof.cuh"

namespace team {

constexpr int TC_BM = 128;           // UMMA_M
constexpr int TC_BK = 64;            // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 128;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a wrong descriptor / byte count traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptors (SWIZZLE_128B, descriptor version 1)
//   K-major : rows of 128 B (64 bf16 along K); 8-row groups SBO = 1024 B apart.
//   MN-major: rows of 128 B (64 bf16 along M/N), one row per k; 8-k groups SBO = 1024 B apart;
//             64-element M/N chunks LBO bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, majors, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;       // 16 KB
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN >= 128) ? 3 : 4;      // 96 KB -> two CTAs per SM overlap epilogue and main loop
    static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// grid = (ceil(N/BN), ceil(M/128), splits); 128 threads.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                         const __grid_constant__ CUtensorMap map_b, int M, int N, int K, int kb_per_split,
                         int has_a2, float alpha, float beta, const float* __restrict__ bias,
                         float* __restrict__ C, int64_t ldc, float* __restrict__ partial) {
    using S = TcSmem<BN>;
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* tmem_full_bar = empty_bar + S::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
    const int nkb = (K + TC_BK - 1) / TC_BK;                 // k-blocks of one pass over K
    const int total_kb = has_a2 ? 2 * nkb : nkb;             // hi pass then lo pass
    const int kb_begin = blockIdx.z * kb_per_split;
    const int kb_end = min(total_kb, kb_begin + kb_per_split);
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (kb_begin < kb_end) {
        if (threadIdx.x == 0) {
            // ===================== TMA producer =====================
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char* sa = smem + stage * S::STAGE_BYTES;
                unsigned char* sb = sa + S::A_BYTES;
                mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                const bool lo = kb >= nkb;
                const CUtensorMap* ma = lo ? &map_a2 : &map_a;
                const int k0 = (lo ? kb - nkb : kb) * TC_BK;
                if (!A_MN) {
                    tma_load_2d(sa, ma, &full_bar[stage], k0, m0);                    // box {64 k, 128 rows}
                } else {
#pragma unroll
                    for (int c = 0; c < TC_BM / 64; ++c)                                // box {64 m, 64 k} per chunk
                        tma_load_2d(sa + c * (TC_BK * 128), ma, &full_bar[stage], m0 + 64 * c, k0);
                }
                if (!B_MN) {
                    tma_load_2d(sb, &map_b, &full_bar[stage], k0, n0);                // box {64 k, BN rows}
                } else {
#pragma unroll
                    for (int c = 0; c < BN / 64; ++c)
                        tma_load_2d(sb + c * (TC_BK * 128), &map_b, &full_bar[stage], n0 + 64 * c, k0);
                }
                if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
            }
        } else if (threadIdx.x == 32) {
            // ===================== MMA issuer (one thread) =====================
            constexpr uint32_t idesc = umma_idesc(TC_BM, BN, A_MN, B_MN);
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb_begin; kb < kb_end; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    const uint64_t ad = A_MN ? umma_desc(sa + k * 2048, TC_BK * 128, 1024)
                                             : umma_desc(sa + k * 32, 0, 1024);
                    const uint64_t bd = B_MN ? umma_desc(sb + k * 2048, TC_BK * 128, 1024)
                                             : umma_desc(sb + k * 32, 0, 1024);
                    tcgen05_mma_f16(tmem_base, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                }
                tcgen05_commit(&empty_bar[stage]);          // smem slot free once these MMAs retire
                if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
            }
            tcgen05_commit(tmem_full_bar);                  // accumulator complete
        }
        // ===================== epilogue: all 4 warps =====================
        __syncwarp();
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int row = m0 + warp * 32 + lane;
    const bool have_acc = kb_begin < kb_end;
    float* P = partial ? partial + (size_t)blockIdx.z * M * N : nullptr;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
        float v[16];
        if (have_acc) {
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);   // warp-collective
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        const int col = n0 + c;
        if (row < M && col < N) {
            if (P != nullptr) {
                float* dst = P + (size_t)row * N + col;
#pragma unroll
                for (int i = 0; i < 16; ++i) if (col + i < N) dst[i] = v[i];
            } else {
                float* dst = C + (int64_t)row * ldc + col;
                const bool vec = (col + 15 < N) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = alpha * v[i];
                    if (bias != nullptr && col + i < N) x += __ldg(bias + col + i);
                    v[i] = x;
                }
                if (vec) {
                    if (beta != 0.f) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 o = *reinterpret_cast<const float4*>(dst + i);
                            v[i] += beta * o.x; v[i + 1] += beta * o.y; v[i + 2] += beta * o.z; v[i + 3] += beta * o.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (col + i < N) dst[i] = v[i] + (beta != 0.f ? beta * dst[i] : 0.f);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

__global__ void __launch_bounds__(256)
tc_splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, float alpha, float beta,
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

// fp32 -> bf16 (hi) and optional residual (lo = bf16(x - hi)); rows x cols with leading dims
__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols,
                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ldd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // index over rows * cols/4
    const int c4 = cols / 4;
    if (i >= rows * c4) return;
    const int64_t r = i / c4;
    const int c = (int)(i % c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(hi + r * ldd + c) = o;
    if (lo != nullptr) {
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
        const __nv_bfloat16 l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
        __nv_bfloat162 c0 = __halves2bfloat162(l0, l1), c1 = __halves2bfloat162(l2, l3);
        o.x = *reinterpret_cast<uint32_t*>(&c0); o.y = *reinterpret_cast<uint32_t*>(&c1);
        *reinterpret_cast<uint2*>(lo + r * ldd + c) = o;
    }
}

// ------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static int get_encode() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    });
    if (g_encode == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return TEAM_EUNSUPPORTED;
    }
    return TEAM_OK;
}

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with leading dimension ld (elements).
static int make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%lld outer=%lld ld=%lld box=%dx%d", (int)r, base,
                  (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer);
        return TEAM_ECUDA;
    }
    return TEAM_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int tc_launch(cudaStream_t st, const TcGemm& g, int splits, int kbps, float* partial) {
    using S = TcSmem<BN>;
    CUtensorMap ma, ma2, mb;
    int rc;
    // K-major operand [rows,K]: inner = K, box {64, tile rows};  MN-major operand [K,rows]: inner = rows, box {64, 64}
    if (!A_MN) rc = make_map(&ma, g.A, g.K, g.M, g.lda, TC_BK, TC_BM); else rc = make_map(&ma, g.A, g.M, g.K, g.lda, 64, TC_BK);
    if (rc) return rc;
    if (g.A2 != nullptr) {
        if (!A_MN) rc = make_map(&ma2, g.A2, g.K, g.M, g.lda, TC_BK, TC_BM); else rc = make_map(&ma2, g.A2, g.M, g.K, g.lda, 64, TC_BK);
        if (rc) return rc;
    } else {
        ma2 = ma;
    }
    if (!B_MN) rc = make_map(&mb, g.B, g.K, g.N, g.ldb, TC_BK, BN); else rc = make_map(&mb, g.B, g.N, g.K, g.ldb, 64, TC_BK);
    if (rc) return rc;
    auto kern = gemm_bf16_tcgen05_kernel<BN, A_MN, B_MN>;
    static bool attr_set = false;          // per template instantiation
    if (!attr_set) {
        TEAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES));
        attr_set = true;
    }
    dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((g.M + TC_BM - 1) / TC_BM), (unsigned)splits);
    const int pslot = prof_enabled() ? prof_begin(st, 1, 2.0 * g.M * g.N * g.K * (g.A2 ? 2 : 1),
                                                  2.0 * (g.M * g.K * (g.A2 ? 2 : 1) + g.N * g.K) + 4.0 * g.M * g.N) : -1;
    kern<<<grid, TC_THREADS, S::BYTES, st>>>(ma, ma2, mb, (int)g.M, (int)g.N, (int)g.K, kbps, g.A2 != nullptr ? 1 : 0,
                                             g.alpha, g.beta, g.bias, g.C, g.ldc, partial);
    if (pslot >= 0) prof_end(st, pslot);
    TEAM_LAUNCH_CHECK("gemm_bf16_tcgen05_kernel");
    if (splits > 1) {
        const int64_t tot = g.M * g.N;
        tc_splitk_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(partial, splits, (int)g.M, (int)g.N, g.alpha, g.beta, g.C, g.ldc, g.bias);
        TEAM_LAUNCH_CHECK("tc_splitk_reduce_kernel");
    }
    return TEAM_OK;
}

int gemm_bf16_tc(cudaStream_t st, const TcGemm& g, void* ws, size_t ws_bytes) {
    int rc = get_encode();
    if (rc) return rc;
    if (g.M <= 0 || g.N <= 0) return TEAM_OK;
    TEAM_REQUIRE(g.K > 0 && g.lda % 8 == 0 && g.ldb % 8 == 0, "gemm_bf16_tc: K=%lld lda=%lld ldb=%lld (leading dims must be multiples of 8)", (long long)g.K, (long long)g.lda, (long long)g.ldb);
    TEAM_REQUIRE((reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0, "gemm_bf16_tc: operands must be 16-byte aligned");
    // tile width: the widest BN that still fills the machine
    const int64_t mt = (g.M + TC_BM - 1) / TC_BM;
    int BN = 128;
    if (g.N <= 64 || mt * ((g.N + 127) / 128) < NUM_SMS) BN = 64;
    const int64_t tiles = mt * ((g.N + BN - 1) / BN);
    const int nkb = (int)((g.K + TC_BK - 1) / TC_BK);
    const int total_kb = g.A2 ? 2 * nkb : nkb;
    int splits = 1;
    if (tiles * 2 <= NUM_SMS && total_kb >= 8 && ws != nullptr) {
        int64_t s = NUM_SMS / tiles;
        if (s > total_kb / 4) s = total_kb / 4;
        if (s >= 2 && (size_t)s * g.M * g.N * sizeof(float) <= ws_bytes) splits = (int)s;
    }
    int kbps = (total_kb + splits - 1) / splits;
    splits = (total_kb + kbps - 1) / kbps;
    float* partial = splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
#define TC_DISPATCH(BN_)                                                                       \
    do {                                                                                       \
        if (!g.a_mn && !g.b_mn) return tc_launch<BN_, false, false>(st, g, splits, kbps, partial); \
        if (!g.a_mn && g.b_mn) return tc_launch<BN_, false, true>(st, g, splits, kbps, partial);   \
        if (g.a_mn && !g.b_mn) return tc_launch<BN_, true, false>(st, g, splits, kbps, partial);   \
        return tc_launch<BN_, true, true>(st, g, splits, kbps, partial);                        \
    } while (0)
    if (BN == 128) TC_DISPATCH(128);
    TC_DISPATCH(64);
#undef TC_DISPATCH
}

int to_bf16(cudaStream_t st, const float* src, int64_t lds, int64_t rows, int cols, void* hi, void* lo, int64_t ldd) {
    TEAM_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "to_bf16: cols/ld must be multiples of 4");
    const int64_t n = rows * (cols / 4);
    if (n == 0) return TEAM_OK;
    f32_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, lds, rows, cols, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), ldd);
    TEAM_LAUNCH_CHECK("f32_to_bf16_kernel");
    return TEAM_OK;
}

size_t tc_operand_bytes(const HeadDims& d) {
    // generous bound on the bf16 copies one fwd or bwd call makes (see BfCache in head.cu)
    const size_t B2 = d.B2, Nsp = d.Nsp;
    const size_t elems = B2 * (20 * (size_t)D + 8 * Nsp) + Nsp * (16 * (size_t)D + 6 * Nsp) + 16 * (size_t)D * D +
                         (size_t)(d.Rt + d.C + 64 + (d.Tc > 0 ? d.Tc : 0)) * 4 * D;
    return align_up(elems * 2 + 96 * 256, 256);
}

}  // namespace team

using namespace team;

extern "C" int team_gemm_bf16(int a_mn, int b_mn, int64_t M, int64_t N, int64_t K, float alpha, const void* A,
                              const void* A_lo, int64_t lda, const void* B, int64_t ldb, float beta, float* C,
                              int64_t ldc, const float* bias, void* workspace, size_t workspace_bytes, void* stream) {
    TcGemm g;
    g.a_mn = a_mn != 0; g.b_mn = b_mn != 0; g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = A; g.A2 = A_lo; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.bias = bias;
    return gemm_bf16_tc((cudaStream_t)stream, g, workspace, workspace_bytes);
}

extern "C" int team_gemm_bf16_nt(int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                                 int64_t ldb, float* C, int64_t ldc, void* stream) {
    return team_gemm_bf16(0, 0, M, N, K, 1.f, A, nullptr, lda, B, ldb, 0.f, C, ldc, nullptr, nullptr, 0, stream);
}

extern "C" int team_f32_to_bf16(const float* src, int64_t lds, int64_t rows, int64_t cols, void* hi, void* lo,
                                int64_t ldd, void* stream) {
    return to_bf16((cudaStream_t)stream, src, lds, rows, (int)cols, hi, lo, ldd);
}
