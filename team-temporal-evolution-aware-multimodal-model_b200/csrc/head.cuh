// Workspace layout + internal declarations of the fusion head (forward_tri_modal fwd/bwd).
// See DESIGN.md section 3 for the algorithm (tests/factorised_model.py is its executable spec).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace team {

struct HeadDims {
    int B;      // samples
    int B2;     // own rows = image rows [0,B) + text rows [B,2B)
    int C;      // classes (prototype rows)
    int P;      // prompt rows
    int M;      // shared key rows = C + P
    int Ns;     // step rows = M + 10 state-table rows
    int Nsp;    // Ns rounded up to 16 (leading dimension of all [.,Ns] matrices)
    int Rt;     // table-query rows = C + 10
    int Tc;     // text-class rows for the classification logits (0 = none)
    int nctas;  // CTAs of the per-sample reduction kernels (size of the partial buffers)
};

// length (floats) of one CTA's partial record written by table_rows_bwd
__host__ __device__ inline size_t table_partial_len(const HeadDims& d) {
    return (size_t)2 * d.Rt * D      // R, G
           + (size_t)d.Rt            // h
           + (size_t)d.Rt * 10       // dTT state columns
           + (size_t)10 * D          // dVF of the state-table rows
           + (size_t)2 * D;          // dgamma, dbeta
}
constexpr int OWN_PARTIAL_LEN = 3 * D;   // dgamma, dbeta, dbfc of the own rows

struct HeadWS {
    // ---- step level
    float *Wsum[3], *bsum[3];         // summed projections img/text/state
    float* Zc;                        // [Tc][D] encode_text(text_cls), pre-normalisation
    float* Ztab;                      // [Rt][D] pre-normalisation proto/state rows
    float* S;                         // [Nsp][D] step rows
    float* invS;                      // [Nsp] inverse norms (proto + state rows)
    float* QKVs;                      // [Nsp][3D]
    float* VFs;                       // [Nsp][D]
    float* TT;                        // [Nsp][Nsp]
    float *mt, *Zt;                   // [Nsp]
    float* Pt;                        // [Nsp][Nsp]
    float* NFt;                       // [Nsp][D]
    // ---- per sample
    float* Xo;                        // [B2][D]
    float* invo;                      // [B2]
    float* QKVo;                      // [B2][3D]
    float* VFo;                       // [B2][D]
    float *SQ, *SK;                   // [B2][Nsp]
    float* Aext;                      // [B2][Nsp]
    float* aown;                      // [B2][2]
    float* Ybo;                       // [B2][D]
    float* lnstat;                    // [B2][2]
    // ---- backward scratch
    float* dYo;                       // [B2][D]
    float* rowdot;                    // [B2]
    float* dsown;                     // [B2][2]
    float* dSK;                       // [B2][Nsp]
    float* dVFo;                      // [B2][D]
    float* dQKVo;                     // [B2][3D]
    float* dXo;                       // [B2][D]
    float *Rfull, *Gfull;             // [Nsp][D]
    float* hfull;                     // [Nsp]
    float* dTT;                       // [Nsp][Nsp]
    float* tmpNN;                     // [Nsp][Nsp]
    float* dVFs;                      // [Nsp][D]
    float* dQKVs;                     // [Nsp][3D]
    float* dZtab;                     // [Rt][D]
    float* tab_partials;              // [nctas][table_partial_len]
    float* tab_reduced;               // [table_partial_len]
    float* own_partials;              // [nctas][OWN_PARTIAL_LEN]
    float* own_reduced;               // [OWN_PARTIAL_LEN]
    float* colsum_partials;           // [64][D]
    void* gemm_ws;                    // split-K scratch
    size_t gemm_ws_bytes;
    // ---- bf16 operands for the tcgen05 path (mode BF16 only)
    void* bf16_area;
    size_t bf16_bytes;
    size_t total_bytes;
};

HeadDims head_dims(int64_t B, int C, int P, int Tc);
// carve `base` (may be null: sizing only) into the HeadWS pointers
void head_plan(const HeadDims& d, int mode, void* base, HeadWS* ws);

int cosine_logits_launch(cudaStream_t st, const float* x, int64_t n_rows, const float* w, int64_t num_classes,
                         const float* sigma_dev, float* logits, int64_t* argmax);

}  // namespace team
