// Graph / state kernels of the TEAM head (K10-K18, a13-a19):
//   temporal GCN over (class, life-stage) prototype nodes   models/dynamic_modal_graph.py:210-337
//   pairwise state distances of the evolved nodes            models/state_evolution.py:345-364
//   state-distance matrix (symmetrise / EMA updates)         utils/state_distance.py:65-144, models/proof.py:666-675
//   class-prototype sync                                     utils/inc_net.py:600-617
//   DynamicGCN weighted scatter-add layers                   models/dynamic_modal_graph.py:131-163
//
// The reference walks edges and node pairs in Python (8 tiny launches per edge, one .item() sync
// per pair).  Here: the per-edge Linear(640->320) is split as W_m = [W_src | W_dst] so that it
// becomes two node-level GEMMs; one warp per destination node then folds its incoming edges
// (dst-sorted CSR, reference edge order) with shuffle-reduced LayerNorms - atomic-free and
// deterministic.  Pairwise distances: one warp per node row, shuffle-reduced dot products,
// per-(state,state) sums kept in double like the reference's Python floats.
#include "head_kernels.cuh"

namespace team {

constexpr int GH = 320;            // hidden 256 + time 64
constexpr int GHID = 256;
constexpr int GT = 64;
constexpr int GPL = GH / 32;       // 10 elements per lane

// LayerNorm over n = 32*PL elements held as v[k] = x[lane + 32k]
template <int PL>
__device__ __forceinline__ void warp_ln(float (&v)[PL], const float* __restrict__ g, const float* __restrict__ b, int lane, bool relu) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < PL; ++k) s += v[k];
    const float mean = warp_sum(s) * (1.0f / (32 * PL));
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < PL; ++k) { v[k] -= mean; q += v[k] * v[k]; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / (32 * PL)) + LN_EPS);
#pragma unroll
    for (int k = 0; k < PL; ++k) {
        float o = v[k] * rstd * g[lane + 32 * k] + b[lane + 32 * k];
        v[k] = relu ? fmaxf(o, 0.f) : o;
    }
}

// in-place L2 normalisation of 512-wide rows (F.normalize, eps 1e-12)
__global__ void __launch_bounds__(256)
graph_rows_normalize_kernel(float* __restrict__ X, int64_t n_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    float4 v[4];
    ld_row(X + r * D, lane, v);
    scale_row(v, 1.0f / fmaxf(sqrtf(warp_sum(dot_part(v, v))), NORM_EPS));
    st_row(X + r * D, lane, v);
}

// z[n] = [ReLU(LN256(Hpre[n])) | ReLU(LN64(t[n]*Wt + bt))]
__global__ void __launch_bounds__(256)
tgcn_encode_kernel(const float* __restrict__ Hpre, const float* __restrict__ time, int n_nodes,
                   const float* __restrict__ ln1_g, const float* __restrict__ ln1_b, const float* __restrict__ Wt,
                   const float* __restrict__ bt, const float* __restrict__ lnt_g, const float* __restrict__ lnt_b,
                   float* __restrict__ z) {
    const int lane = threadIdx.x & 31, n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= n_nodes) return;
    float h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = Hpre[(size_t)n * GHID + lane + 32 * k];
    warp_ln<8>(h, ln1_g, ln1_b, lane, true);
#pragma unroll
    for (int k = 0; k < 8; ++k) z[(size_t)n * GH + lane + 32 * k] = h[k];
    const float t = time[n];
    float tv[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) tv[k] = t * Wt[lane + 32 * k] + bt[lane + 32 * k];
    warp_ln<2>(tv, lnt_g, lnt_b, lane, true);
#pragma unroll
    for (int k = 0; k < 2; ++k) z[(size_t)n * GH + GHID + lane + 32 * k] = tv[k];
}

// M[d] = (sum_{e: dst=d} w_e * ReLU(LN320(Us[src_e] + Ud[d]))) / (deg + 1e-8)   (0 if deg == 0)
__global__ void __launch_bounds__(256)
tgcn_message_kernel(const float* __restrict__ Us, const float* __restrict__ Ud, const int* __restrict__ rowptr,
                    const int* __restrict__ src, const float* __restrict__ ew, int n_nodes,
                    const float* __restrict__ ln_g, const float* __restrict__ ln_b, float* __restrict__ Mout) {
    const int lane = threadIdx.x & 31, d = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (d >= n_nodes) return;
    float ud[GPL], acc[GPL];
#pragma unroll
    for (int k = 0; k < GPL; ++k) { ud[k] = Ud[(size_t)d * GH + lane + 32 * k]; acc[k] = 0.f; }
    const int e0 = rowptr[d], e1 = rowptr[d + 1];
    for (int e = e0; e < e1; ++e) {
        const int s = src[e];
        const float w = ew[e];
        float v[GPL];
#pragma unroll
        for (int k = 0; k < GPL; ++k) v[k] = Us[(size_t)s * GH + lane + 32 * k] + ud[k];
        warp_ln<GPL>(v, ln_g, ln_b, lane, true);
#pragma unroll
        for (int k = 0; k < GPL; ++k) acc[k] += v[k] * w;
    }
    const float cnt = (float)(e1 - e0);
    const float sc = e1 > e0 ? 1.0f / (cnt + 1e-8f) : 0.f;
#pragma unroll
    for (int k = 0; k < GPL; ++k) Mout[(size_t)d * GH + lane + 32 * k] = acc[k] * sc;
}

// z_out = g * ReLU(LN320(Tpre)) + (1-g) * z,  g = sigmoid(<wg, z> + bg)
__global__ void __launch_bounds__(256)
tgcn_update_kernel(const float* __restrict__ Tpre, const float* __restrict__ z, int n_nodes,
                   const float* __restrict__ wg, const float* __restrict__ bg, const float* __restrict__ ln_g,
                   const float* __restrict__ ln_b, float* __restrict__ zout) {
    const int lane = threadIdx.x & 31, n = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= n_nodes) return;
    float zv[GPL], u[GPL];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < GPL; ++k) {
        zv[k] = z[(size_t)n * GH + lane + 32 * k];
        u[k] = Tpre[(size_t)n * GH + lane + 32 * k];
        dot += zv[k] * wg[lane + 32 * k];
    }
    const float gate = 1.0f / (1.0f + expf(-(warp_sum(dot) + bg[0])));
    warp_ln<GPL>(u, ln_g, ln_b, lane, true);
#pragma unroll
    for (int k = 0; k < GPL; ++k) zout[(size_t)n * GH + lane + 32 * k] = gate * u[k] + (1.0f - gate) * zv[k];
}

// one warp per node row i: d_ij = 1 - cos(u_i,u_j) for all j != i, summed per state of j in double
__global__ void __launch_bounds__(256)
pairwise_dist_kernel(const float* __restrict__ U, const int* __restrict__ states, int n_nodes,
                     double* __restrict__ rowsum, int* __restrict__ rowcnt) {
    __shared__ double acc[8][10];
    __shared__ int cnt[8][10];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * 8 + w;
    if (lane < 10) { acc[w][lane] = 0.0; cnt[w][lane] = 0; }
    __syncwarp();
    if (i >= n_nodes) return;
    float4 ui[4], uj[4];
    ld_row(U + (size_t)i * D, lane, ui);
    const float ni = sqrtf(warp_sum(dot_part(ui, ui)));
    for (int j = 0; j < n_nodes; ++j) {
        if (j == i) continue;
        ld_row(U + (size_t)j * D, lane, uj);
        const float dt = warp_sum(dot_part(ui, uj));
        const float nj = sqrtf(warp_sum(dot_part(uj, uj)));
        // F.cosine_similarity: x.y / (max(|x|,eps) * max(|y|,eps)), eps = 1e-8
        const float sim = dt / (fmaxf(ni, 1e-8f) * fmaxf(nj, 1e-8f));
        if (lane == 0) {
            const int sj = states[j];
            acc[w][sj] += 1.0 - (double)sim;
            cnt[w][sj] += 1;
        }
    }
    __syncwarp();
    if (lane < 10) { rowsum[(size_t)i * 10 + lane] = acc[w][lane]; rowcnt[(size_t)i * 10 + lane] = cnt[w][lane]; }
}

// sums[a][b] = sum over rows i with state a (ascending i) of rowsum[i][b]
__global__ void pairwise_reduce_kernel(const double* __restrict__ rowsum, const int* __restrict__ rowcnt,
                                       const int* __restrict__ states, int n_nodes, double* __restrict__ sums,
                                       long long* __restrict__ counts) {
    const int t = threadIdx.x;
    if (t >= 100) return;
    const int a = t / 10, b = t % 10;
    double s = 0.0;
    long long c = 0;
    for (int i = 0; i < n_nodes; ++i)
        if (states[i] == a) { s += rowsum[(size_t)i * 10 + b]; c += rowcnt[(size_t)i * 10 + b]; }
    sums[t] = s;
    counts[t] = c;
}

// P_c = normalize(sum_i (w_i / sum w) p_i),  w = 1.5 for state 4 else 1   (one warp per class group)
__global__ void __launch_bounds__(256)
sync_protos_kernel(const float* __restrict__ nodes, const int* __restrict__ grp_ptr, const int* __restrict__ node_state,
                   const int* __restrict__ grp_class, int n_groups, float* __restrict__ protos) {
    const int lane = threadIdx.x & 31, g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const int i0 = grp_ptr[g], i1 = grp_ptr[g + 1];
    float wsum = 0.f;
    for (int i = i0; i < i1; ++i) wsum += node_state[i] == 4 ? 1.5f : 1.0f;
    float4 acc[4], p[4];
    zero_row(acc);
    for (int i = i0; i < i1; ++i) {
        const float w = (node_state[i] == 4 ? 1.5f : 1.0f) / wsum;
        ld_row(nodes + (size_t)i * D, lane, p);
        axpy_row(acc, w, p);
    }
    const float inv = 1.0f / fmaxf(sqrtf(warp_sum(dot_part(acc, acc))), NORM_EPS);
    scale_row(acc, inv);
    st_row(protos + (size_t)grp_class[g] * D, lane, acc);
}


// out[g] = (sum_{i in group g} nodes[member[i]]) / n_g   (torch.stack(rows).mean(0); one warp per group)
__global__ void __launch_bounds__(256)
group_mean_kernel(const float* __restrict__ nodes, const int* __restrict__ grp_ptr, const int* __restrict__ member,
                  int n_groups, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const int i0 = grp_ptr[g], i1 = grp_ptr[g + 1];
    float4 acc[4], p[4];
    zero_row(acc);
    for (int i = i0; i < i1; ++i) {
        ld_row(nodes + (size_t)(member ? member[i] : i) * D, lane, p);
        add_row(acc, p);
    }
    const float n = (float)(i1 - i0);
    if (i1 > i0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[k].x /= n; acc[k].y /= n; acc[k].z /= n; acc[k].w /= n; }
    }
    st_row(out + (size_t)g * D, lane, acc);
}

// out = (F + F^T)/2 off the diagonal, 1 on it
__global__ void dist_matrix_kernel(const float* __restrict__ F, int n, float* __restrict__ out) {
    const int t = threadIdx.x;
    if (t >= n * n) return;
    const int i = t / n, j = t % n;
    out[t] = i == j ? 1.0f : (F[i * n + j] + F[j * n + i]) / 2.0f;
}

// sequential EMA with the reference's double-visit semantics (models/proof.py:666-675):
//   for (s1,s2,d) in order: F[s1,s2] = F[s2,s1] = (1-w)*F[s1,s2] + w*d      (python-float arithmetic)
__global__ void dist_ema_kernel(float* __restrict__ F, int n, const int* __restrict__ keys, const double* __restrict__ vals,
                                int m, double weight) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int e = 0; e < m; ++e) {
        const int s1 = keys[2 * e], s2 = keys[2 * e + 1];
        const double old = (double)F[s1 * n + s2];
        const float nw = (float)((1.0 - weight) * old + weight * vals[e]);
        F[s1 * n + s2] = nw;
        F[s2 * n + s1] = nw;
    }
}

// AdaptiveStateDistanceMatrix.forward update branch (utils/state_distance.py:96-134): centres of the
// states 1..9 present in the batch -> 2 - cosine -> sequential EMA.  One CTA, 10 warps.
__global__ void __launch_bounds__(320)
state_dist_forward_kernel(const float* __restrict__ sums, const long long* __restrict__ counts, float* __restrict__ F,
                          double decay) {
    __shared__ __align__(16) float cen[10][D];
    __shared__ float sim[10][10];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool present = w >= 1 && counts[w] > 0;
    float4 c[4];
    zero_row(c);
    if (present) {
        ld_row(sums + (size_t)w * D, lane, c);
        const float fn = (float)counts[w];
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[i].x /= fn; c[i].y /= fn; c[i].z /= fn; c[i].w /= fn; }
        const float inv = 1.0f / fmaxf(sqrtf(warp_sum(dot_part(c, c))), NORM_EPS);
        scale_row(c, inv);
    }
    st_row(cen[w], lane, c);
    __syncthreads();
    for (int j = 0; j < 10; ++j) {
        float4 o[4];
        ld_row(cen[j], lane, o);
        const float dt = warp_sum(dot_part(c, o));
        if (lane == 0) sim[w][j] = dt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ids[10], n = 0;
        for (int s = 1; s < 10; ++s) if (counts[s] > 0) ids[n++] = s;
        if (n > 1) {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    if (i == j) continue;
                    const int si = ids[i], sj = ids[j];
                    const double old = (double)F[si * 10 + sj];
                    const double dm = (double)(2.0f - sim[si][sj]);
                    const float nw = (float)(decay * old + (1.0 - decay) * dm);
                    F[si * 10 + sj] = nw;
                    F[sj * 10 + si] = nw;
                }
        }
    }
}

// DynamicGCN layer tail: hu[d] = relu(h[d]) + sum_{e: dst=d} w_e relu(h[src_e]);  out = LN(hu)
__global__ void __launch_bounds__(256)
dgcn_aggregate_ln_kernel(const float* __restrict__ Hpre, int n_nodes, int dim, const int* __restrict__ rowptr,
                         const int* __restrict__ src, const float* __restrict__ ew, const float* __restrict__ ln_g,
                         const float* __restrict__ ln_b, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, d = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (d >= n_nodes) return;
    const int pl = dim / 32;          // <= 16
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = k < pl ? fmaxf(Hpre[(size_t)d * dim + lane + 32 * k], 0.f) : 0.f;
    if (rowptr != nullptr) {
        for (int e = rowptr[d]; e < rowptr[d + 1]; ++e) {
            const int s = src[e];
            const float w = ew[e];
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k < pl) v[k] += w * fmaxf(Hpre[(size_t)s * dim + lane + 32 * k], 0.f);
        }
    }
    float sm = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) sm += v[k];
    const float mean = warp_sum(sm) / (float)dim;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) if (k < pl) { v[k] -= mean; q += v[k] * v[k]; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)dim + LN_EPS);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < pl) out[(size_t)d * dim + lane + 32 * k] = v[k] * rstd * ln_g[lane + 32 * k] + ln_b[lane + 32 * k];
}

}  // namespace team

using namespace team;

#define GG(...)                                                                          \
    do {                                                                                 \
        int _rc = gemm_f32(st, __VA_ARGS__, nullptr, 0);                                 \
        if (_rc != TEAM_OK) return _rc;                                                  \
    } while (0)

extern "C" size_t team_tgcn_workspace_bytes(int64_t n_nodes) {
    // Hpre[N,256] z0[N,320] z1[N,320] Us[N,320] Ud[N,320] M[N,320] T[N,320] O[N,512]
    return align_up((size_t)n_nodes * (GHID + 6 * GH + D) * sizeof(float), 256) + 256;
}

extern "C" int team_tgcn_forward(const team_tgcn_weights* tw, const float* node_feat, const float* time_steps,
                                 int64_t n_nodes, const int32_t* rowptr, const int32_t* src,
                                 const float* edge_w, float* out, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    TEAM_REQUIRE(tw && node_feat && time_steps && rowptr && out && n_nodes >= 1, "team_tgcn_forward: bad args");
    TEAM_REQUIRE(tw->num_blocks >= 1 && tw->num_blocks <= 8, "team_tgcn_forward: num_blocks %d", tw->num_blocks);
    if (workspace == nullptr || workspace_bytes < team_tgcn_workspace_bytes(n_nodes)) {
        set_error("team_tgcn_forward: workspace too small");
        return TEAM_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = n_nodes;
    float* Hpre = reinterpret_cast<float*>(workspace);
    float* z0 = Hpre + N * GHID;
    float* z1 = z0 + N * GH;
    float* Us = z1 + N * GH;
    float* Ud = Us + N * GH;
    float* Mb = Ud + N * GH;
    float* Tb = Mb + N * GH;
    const unsigned wg = (unsigned)((N + 7) / 8);
    GG(false, true, N, GHID, D, 1.f, node_feat, D, tw->node_w, D, 0.f, Hpre, GHID, tw->node_b);
    tgcn_encode_kernel<<<wg, 256, 0, st>>>(Hpre, time_steps, (int)N, tw->node_ln_g, tw->node_ln_b, tw->time_w, tw->time_b, tw->time_ln_g, tw->time_ln_b, z0);
    TEAM_LAUNCH_CHECK("tgcn_encode_kernel");
    float* zin = z0;
    float* zout = z1;
    for (int b = 0; b < tw->num_blocks; ++b) {
        const team_tgcn_block& bk = tw->blocks[b];
        // message_net Linear(640->320) split into source and destination halves (node-level GEMMs)
        GG(false, true, N, GH, GH, 1.f, zin, GH, bk.msg_w, 2 * GH, 0.f, Us, GH, nullptr);
        GG(false, true, N, GH, GH, 1.f, zin, GH, bk.msg_w + GH, 2 * GH, 0.f, Ud, GH, bk.msg_b);
        tgcn_message_kernel<<<wg, 256, 0, st>>>(Us, Ud, rowptr, src, edge_w, (int)N, bk.msg_ln_g, bk.msg_ln_b, Mb);
        TEAM_LAUNCH_CHECK("tgcn_message_kernel");
        GG(false, true, N, GH, GH, 1.f, zin, GH, bk.upd_w, 2 * GH, 0.f, Tb, GH, bk.upd_b);
        GG(false, true, N, GH, GH, 1.f, Mb, GH, bk.upd_w + GH, 2 * GH, 1.f, Tb, GH, nullptr);
        tgcn_update_kernel<<<wg, 256, 0, st>>>(Tb, zin, (int)N, bk.gate_w, bk.gate_b, bk.upd_ln_g, bk.upd_ln_b, zout);
        TEAM_LAUNCH_CHECK("tgcn_update_kernel");
        float* t = zin; zin = zout; zout = t;
    }
    GG(false, true, N, D, GH, 1.f, zin, GH, tw->out_w, GH, 0.f, out, D, tw->out_b);
    graph_rows_normalize_kernel<<<wg, 256, 0, st>>>(out, N);
    TEAM_LAUNCH_CHECK("graph_rows_normalize_kernel");
    return TEAM_OK;
}

// in-place F.normalize of [n,512] rows (evolve_state_prototypes re-normalisation, utils/inc_net.py:595)
extern "C" int team_rows_normalize(float* x, int64_t n_rows, void* stream) {
    TEAM_REQUIRE(x != nullptr && n_rows >= 0, "team_rows_normalize: bad args");
    if (n_rows == 0) return TEAM_OK;
    graph_rows_normalize_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, n_rows);
    TEAM_LAUNCH_CHECK("graph_rows_normalize_kernel");
    return TEAM_OK;
}

extern "C" int team_pairwise_state_dist(const float* node_feat, const int32_t* node_states, int64_t n_nodes,
                                        double* sums, int64_t* counts, void* workspace, size_t workspace_bytes,
                                        void* stream) {
    TEAM_REQUIRE(node_feat && node_states && sums && counts && n_nodes >= 1, "team_pairwise_state_dist: bad args");
    const size_t need = (size_t)n_nodes * 10 * (sizeof(double) + sizeof(int)) + 256;
    if (workspace == nullptr || workspace_bytes < need) { set_error("team_pairwise_state_dist: workspace %zu < %zu", workspace_bytes, need); return TEAM_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    double* rowsum = reinterpret_cast<double*>(workspace);
    int* rowcnt = reinterpret_cast<int*>(rowsum + n_nodes * 10);
    pairwise_dist_kernel<<<(unsigned)((n_nodes + 7) / 8), 256, 0, st>>>(node_feat, node_states, (int)n_nodes, rowsum, rowcnt);
    TEAM_LAUNCH_CHECK("pairwise_dist_kernel");
    pairwise_reduce_kernel<<<1, 128, 0, st>>>(rowsum, rowcnt, node_states, (int)n_nodes, sums, reinterpret_cast<long long*>(counts));
    TEAM_LAUNCH_CHECK("pairwise_reduce_kernel");
    return TEAM_OK;
}

extern "C" int team_sync_prototypes(const float* nodes, const int32_t* group_ptr, const int32_t* node_states,
                                    const int32_t* group_class, int64_t n_groups, float* img_prototypes, void* stream) {
    TEAM_REQUIRE(nodes && group_ptr && node_states && group_class && img_prototypes, "team_sync_prototypes: bad args");
    if (n_groups <= 0) return TEAM_OK;
    sync_protos_kernel<<<(unsigned)((n_groups + 7) / 8), 256, 0, (cudaStream_t)stream>>>(nodes, group_ptr, node_states, group_class, (int)n_groups, img_prototypes);
    TEAM_LAUNCH_CHECK("sync_protos_kernel");
    return TEAM_OK;
}


extern "C" int team_group_mean(const float* nodes, const int32_t* group_ptr, const int32_t* member, int64_t n_groups,
                               float* out, void* stream) {
    TEAM_REQUIRE(nodes && group_ptr && out, "team_group_mean: bad args");
    if (n_groups <= 0) return TEAM_OK;
    group_mean_kernel<<<(unsigned)((n_groups + 7) / 8), 256, 0, (cudaStream_t)stream>>>(nodes, group_ptr, member, (int)n_groups, out);
    TEAM_LAUNCH_CHECK("group_mean_kernel");
    return TEAM_OK;
}

extern "C" int team_dist_matrix(const float* factors, int32_t n, float* out, void* stream) {
    TEAM_REQUIRE(factors && out && n >= 1 && n <= 16, "team_dist_matrix: bad args");
    dist_matrix_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(factors, n, out);
    TEAM_LAUNCH_CHECK("dist_matrix_kernel");
    return TEAM_OK;
}

extern "C" int team_dist_ema(float* factors, int32_t n, const int32_t* keys, const double* vals, int32_t m,
                             double weight, void* stream) {
    TEAM_REQUIRE(factors && n >= 1 && (m == 0 || (keys && vals)), "team_dist_ema: bad args");
    if (m == 0) return TEAM_OK;
    dist_ema_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(factors, n, keys, vals, m, weight);
    TEAM_LAUNCH_CHECK("dist_ema_kernel");
    return TEAM_OK;
}

extern "C" int team_state_dist_forward(const float* state_sums, const int64_t* state_counts, float* factors,
                                       double decay, float* pre_update_matrix, void* stream) {
    TEAM_REQUIRE(state_sums && state_counts && factors, "team_state_dist_forward: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (pre_update_matrix != nullptr) {
        dist_matrix_kernel<<<1, 256, 0, st>>>(factors, 10, pre_update_matrix);
        TEAM_LAUNCH_CHECK("dist_matrix_kernel");
    }
    state_dist_forward_kernel<<<1, 320, 0, st>>>(state_sums, reinterpret_cast<const long long*>(state_counts), factors, decay);
    TEAM_LAUNCH_CHECK("state_dist_forward_kernel");
    return TEAM_OK;
}

extern "C" size_t team_dgcn_workspace_bytes(int64_t n_nodes, int32_t max_dim) {
    return align_up((size_t)n_nodes * max_dim * sizeof(float) * 3, 256) + 256;
}

extern "C" int team_dgcn_forward(const team_dgcn_layer* layers, int32_t n_layers, const float* x, int64_t n_nodes,
                                 const int32_t* rowptr, const int32_t* src, const float* edge_w, float* out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    TEAM_REQUIRE(layers && n_layers >= 1 && x && out && n_nodes >= 1, "team_dgcn_forward: bad args");
    int maxd = 0;
    for (int l = 0; l < n_layers; ++l) {
        TEAM_REQUIRE(layers[l].out_dim % 32 == 0 && layers[l].out_dim <= 512 && layers[l].in_dim >= 1, "team_dgcn_forward: layer %d dims", l);
        if (layers[l].out_dim > maxd) maxd = layers[l].out_dim;
    }
    if (workspace == nullptr || workspace_bytes < team_dgcn_workspace_bytes(n_nodes, maxd)) { set_error("team_dgcn_forward: workspace too small"); return TEAM_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    float* Hpre = reinterpret_cast<float*>(workspace);
    float* buf[2] = {Hpre + n_nodes * maxd, Hpre + 2 * n_nodes * maxd};
    const float* cur = x;
    for (int l = 0; l < n_layers; ++l) {
        const team_dgcn_layer& L = layers[l];
        GG(false, true, n_nodes, L.out_dim, L.in_dim, 1.f, cur, L.in_dim, L.w, L.in_dim, 0.f, Hpre, L.out_dim, L.b);
        float* dst = (l == n_layers - 1) ? out : buf[l & 1];
        dgcn_aggregate_ln_kernel<<<(unsigned)((n_nodes + 7) / 8), 256, 0, st>>>(Hpre, (int)n_nodes, L.out_dim, rowptr, src, edge_w, L.ln_g, L.ln_b, dst);
        TEAM_LAUNCH_CHECK("dgcn_aggregate_ln_kernel");
        cur = dst;
    }
    return TEAM_OK;
}
