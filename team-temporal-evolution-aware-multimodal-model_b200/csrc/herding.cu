// Exemplar herding on the GPU (second half).
// Replaces the selection loop of BaseLearner._construct_exemplar (models/base.py:284-311): for every class,
//   vectors = features / (||features|| + 1e-8);  class_mean = mean(vectors, 0)
//   for k = 1..m:  S = sum of the exemplars chosen so far;  mu_p = (vectors + S) / k
//                  i = argmin_i sqrt(sum_c (class_mean - mu_p)^2)  over the rows not chosen yet (first index on ties)
// and the exemplar mean of :335-341 (mean of the chosen normalised rows, divided by its norm).
// One CTA per class (the classes of a task are independent), 1024 threads: the m picks are a serial chain of global
// argmins over the class's rows, each pick one pass over rows that stay L2-resident (n x 2 KB per class).
// Arithmetic follows the reference's float32 numpy expressions term by term ((v + S) / k, difference, square, sqrt); the
// column sums of class_mean and S run in the reference's order (row by row / in pick order).  Only the 512-term row sum of
// the squared differences has another association (warp tree instead of numpy's pairwise blocks), so a pick can differ
// from numpy only where two candidates' distances agree to fp32 round-off.
#include "common.cuh"

namespace team {

constexpr int HD_THREADS = 1024;
constexpr int HD_WARPS = HD_THREADS / 32;
constexpr float HD_EPS = 1e-8f;            // EPSILON of models/base.py:12

__global__ void __launch_bounds__(HD_THREADS, 1)
herding_kernel(const float* __restrict__ feats, const int64_t* __restrict__ group_ptr, int m, float* __restrict__ vn,
               unsigned char* __restrict__ taken, int64_t* __restrict__ out_idx, float* __restrict__ out_mean,
               float* __restrict__ out_class_mean) {
    __shared__ float cmean[D];
    __shared__ float S[D];
    __shared__ float emean[D];
    __shared__ float best_d[HD_WARPS];
    __shared__ int best_i[HD_WARPS];
    __shared__ int pick;
    __shared__ float red[HD_WARPS];
    pdl_trigger();
    pdl_wait();
    const int g = blockIdx.x;
    const int64_t r0 = group_ptr[g];
    const int n = (int)(group_ptr[g + 1] - r0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* X = feats + r0 * D;
    float* V = vn + r0 * D;
    unsigned char* tk = taken + r0;
    // ---- normalise: v / (||v|| + eps)
    for (int i = warp; i < n; i += HD_WARPS) {
        float4 x[4];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j] = *reinterpret_cast<const float4*>(X + (size_t)i * D + 4 * (lane + 32 * j));
            ss += x[j].x * x[j].x + x[j].y * x[j].y + x[j].z * x[j].z + x[j].w * x[j].w;
        }
        const float nrm = sqrtf(warp_sum(ss)) + HD_EPS;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(V + (size_t)i * D + 4 * (lane + 32 * j)) =
                make_float4(__fdiv_rn(x[j].x, nrm), __fdiv_rn(x[j].y, nrm), __fdiv_rn(x[j].z, nrm), __fdiv_rn(x[j].w, nrm));
        if (lane == 0) tk[i] = 0;
    }
    __syncthreads();
    // ---- class mean: column sums row by row (numpy's add.reduce order along axis 0), then / n
    if (tid < D) {
        float s = 0.f;
        int i = 0;
        for (; i + 8 <= n; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = V[(size_t)(i + u) * D + tid];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; i < n; ++i) s += V[(size_t)i * D + tid];
        cmean[tid] = __fdiv_rn(s, (float)n);
        S[tid] = 0.f;
        emean[tid] = 0.f;
        if (out_class_mean != nullptr) out_class_mean[(size_t)g * D + tid] = cmean[tid];
    }
    __syncthreads();
    // ---- m picks
    for (int k = 1; k <= m; ++k) {
        const float fk = (float)k;
        float4 cm[4], sv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cm[j] = *reinterpret_cast<const float4*>(cmean + 4 * (lane + 32 * j));
            sv[j] = *reinterpret_cast<const float4*>(S + 4 * (lane + 32 * j));
        }
        float bd = INFINITY;
        int bi = 0x7fffffff;
        for (int i = warp; i < n; i += HD_WARPS) {
            if (tk[i]) continue;
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(V + (size_t)i * D + 4 * (lane + 32 * j));
                const float dx = cm[j].x - __fdiv_rn(v.x + sv[j].x, fk), dy = cm[j].y - __fdiv_rn(v.y + sv[j].y, fk);
                const float dz = cm[j].z - __fdiv_rn(v.z + sv[j].z, fk), dw = cm[j].w - __fdiv_rn(v.w + sv[j].w, fk);
                ss += dx * dx + dy * dy + dz * dz + dw * dw;
            }
            const float dist = sqrtf(warp_sum(ss));
            if (dist < bd) { bd = dist; bi = i; }                // rows ascend within a warp: '<' keeps the first index
        }
        if (lane == 0) { best_d[warp] = bd; best_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            float d = best_d[lane];
            int i = best_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
                const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
                if (d2 < d || (d2 == d && i2 < i)) { d = d2; i = i2; }
            }
            if (lane == 0) {                                      // fewer rows than picks: -1 (the reference raises there)
                const bool ok = i != 0x7fffffff;
                pick = ok ? i : -1;
                if (ok) tk[i] = 1;
                out_idx[(size_t)g * m + (k - 1)] = ok ? i : -1;
            }
        }
        __syncthreads();
        if (tid < D && pick >= 0) {
            const float v = V[(size_t)pick * D + tid];
            S[tid] += v;                                          // np.sum(exemplar_vectors, axis=0): pick order
            emean[tid] += v;
        }
        __syncthreads();
    }
    // ---- exemplar mean (models/base.py:338-340): mean of the chosen normalised rows, divided by its norm
    float sq = 0.f;
    if (tid < D) {
        emean[tid] = __fdiv_rn(emean[tid], (float)m);
        sq = emean[tid] * emean[tid];
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    if (tid < D) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < D / 32; ++w) tot += red[w];
        out_mean[(size_t)g * D + tid] = __fdiv_rn(emean[tid], sqrtf(tot));
    }
}

}  // namespace team

using namespace team;

extern "C" size_t team_herding_workspace_bytes(int64_t n_rows) {
    return align_up((size_t)(n_rows > 0 ? n_rows : 1) * D * sizeof(float), 256) + align_up((size_t)(n_rows > 0 ? n_rows : 1), 256);
}

extern "C" int team_herding_select(const float* feats, const int64_t* group_ptr, int32_t n_groups, int32_t m, int64_t n_rows,
                                   int64_t* out_idx, float* out_mean, float* out_class_mean, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    TEAM_REQUIRE(feats && group_ptr && out_idx && out_mean && n_groups >= 1 && m >= 1 && n_rows >= 1, "herding: bad arguments");
    TEAM_REQUIRE(workspace != nullptr && workspace_bytes >= team_herding_workspace_bytes(n_rows), "herding: workspace too small");
    TEAM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "herding: workspace must be 256-byte aligned");
    float* vn = reinterpret_cast<float*>(workspace);
    unsigned char* taken = reinterpret_cast<unsigned char*>(workspace) + align_up((size_t)n_rows * D * sizeof(float), 256);
    TEAM_LAUNCH(herding_kernel, n_groups, HD_THREADS, 0, (cudaStream_t)stream, feats, group_ptr, (int)m, vn, taken, out_idx, out_mean, out_class_mean);
    return TEAM_OK;
}
