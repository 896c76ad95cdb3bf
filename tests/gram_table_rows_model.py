"""Executable specification (fp64, CPU) of the GRAM formulation of the table-query rows - the next step for the
kernels `table_rows_fwd2 / table_rows_bwd2`.  Test helper, not the product.

For sample b and table row tr (C prototype rows, then the state-table row of the sample) the head forms
    u = cw n_tr + s_tr + ai vi_b + at vt_b + as vs_b        (n = NF row, s = S row + b_fc, v* = VF rows)
    xhat = LayerNorm-normalised u,   out_proto[b] = gamma (1/C) sum_{tr<C} xhat + beta,   out_state[b] = gamma xhat_state + beta
and the backward needs dY = d loss / d u and a handful of its dot products and weighted sums.  u, xhat and dY are all
linear combinations of SEVEN vectors {n_tr, s_tr, vi_b, vt_b, vs_b, gg_b, 1} with scalar coefficients, so
  * every scalar (mean, variance, m2, the score gradients dY.v) follows from the Gram matrix of those vectors:
      table x table   (per step:   O(Rt) dots),
      sample x table  (ONE GEMM:   [vi; vt; gg_p; gg_s] (4B rows) x [n; s; vs-table]^T),
      sample x sample (per sample: 7 dots and 4 sums),
  * every vector output (out_proto, sum_k ai dY, ...) is  coefficients x table rows  (a GEMM with K = 2 Rt) plus a
    few per-sample scalars times the sample's own vectors.
`direct` is what the kernels do today (512-wide arithmetic per (sample, row)); `gram` never touches a 512-wide vector
per (sample, row).  tests/test_gram_table_rows.py checks that the two agree to fp64 round-off.
"""
from __future__ import annotations

import torch

LN_EPS = 1e-5


def make_case(B=6, C=4, D=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(s, generator=g, dtype=torch.float64)
    Rt = C + 10
    case = {"N": rn(Rt, D), "S": rn(Rt, D), "VI": rn(B, D), "VT": rn(B, D), "VStab": rn(10, D),
            "gamma": 1 + 0.1 * rn(D), "beta": 0.1 * rn(D), "g_proto": rn(B, D), "g_state": rn(B, D),
            "sid": torch.randint(0, 10, (B,), generator=g)}
    coef = torch.rand((B, C + 1, 4), generator=g, dtype=torch.float64)
    case["coef"] = coef / coef.sum(-1, keepdim=True) * 0.9           # (cw, ai, at, as): attention weights
    return case


def rows_of(c, b):
    C = c["coef"].shape[1] - 1
    return list(range(C)) + [C + int(c["sid"][b])]


def direct(c):
    """512-wide arithmetic per (sample, row): the reference semantics the kernels implement today."""
    B, D = c["VI"].shape
    C = c["coef"].shape[1] - 1
    out_p, out_s = torch.zeros(B, D, dtype=torch.float64), torch.zeros(B, D, dtype=torch.float64)
    dVI, dVT, dVS = torch.zeros_like(out_p), torch.zeros_like(out_p), torch.zeros_like(out_p)
    kv = torch.zeros(B, C + 1, 4, dtype=torch.float64)                # dY.ybar, dY.vi, dY.vt, dY.vs
    for b in range(B):
        vs = c["VStab"][int(c["sid"][b])]
        for k, tr in enumerate(rows_of(c, b)):
            cw, ai, at, as_ = c["coef"][b, k]
            ybar = cw * c["N"][tr] + ai * c["VI"][b] + at * c["VT"][b] + as_ * vs
            u = ybar + c["S"][tr]
            mean = u.mean(); var = ((u - mean) ** 2).mean(); rstd = (var + LN_EPS).rsqrt()
            xh = (u - mean) * rstd
            if k < C:
                out_p[b] += xh / C
                gg = c["gamma"] * c["g_proto"][b] / C
            else:
                out_s[b] = xh
                gg = c["gamma"] * c["g_state"][b]
            dY = rstd * (gg - gg.mean() - xh * (gg * xh).mean())
            dVI[b] += ai * dY; dVT[b] += at * dY; dVS[b] += as_ * dY
            kv[b, k] = torch.stack([dY @ ybar, dY @ c["VI"][b], dY @ c["VT"][b], dY @ vs])
    return {"out_proto": out_p * c["gamma"] + c["beta"], "out_state": out_s * c["gamma"] + c["beta"],
            "dVI": dVI, "dVT": dVT, "dVS": dVS, "kv": kv}


def gram(c):
    """The same outputs from Gram entries and coefficient GEMMs only."""
    B, D = c["VI"].shape
    C = c["coef"].shape[1] - 1
    Rt = C + 10
    N, S, VI, VT, VSt = c["N"], c["S"], c["VI"], c["VT"], c["VStab"]
    GGp, GGs = c["gamma"] * c["g_proto"] / C, c["gamma"] * c["g_state"]
    # ---- Gram entries
    Tb = torch.cat([N, S, VSt], 0)                                    # [2Rt + 10, D]  table side of the one GEMM
    A = torch.cat([VI, VT, GGp, GGs], 0)                              # [4B, D]        sample side
    ST = (A @ Tb.t()).view(4, B, 2 * Rt + 10)                         # sample x table
    TTm = Tb @ Tb.t()                                                 # table x table (tiny, per step)
    sumT = Tb.sum(1)
    sums = A.sum(1).view(4, B)                                        # sum vi, sum vt, sum gg_p, sum gg_s
    SS = {(i, j): (A.view(4, B, D)[i] * A.view(4, B, D)[j]).sum(1) for i in range(4) for j in range(4)}   # sample x sample
    # ---- per (b, k) scalars, then coefficient matrices
    cf_out = torch.zeros(B, 2 * Rt, dtype=torch.float64)              # out_proto: table part
    sc_out = torch.zeros(B, 4, dtype=torch.float64)                   # x vi, vt, vs, 1
    out_s = torch.zeros(B, D, dtype=torch.float64)
    cf_d = torch.zeros(3, B, 2 * Rt, dtype=torch.float64)             # dVI / dVT / dVS: table part
    sc_d = torch.zeros(3, B, 6, dtype=torch.float64)                  # x gg_p, gg_s, vi, vt, vs, 1
    kv = torch.zeros(B, C + 1, 4, dtype=torch.float64)
    for b in range(B):
        st = int(c["sid"][b])
        iv = 2 * Rt + st                                              # column of vs in the table side
        for k, tr in enumerate(rows_of(c, b)):
            cw, ai, at, as_ = [float(x) for x in c["coef"][b, k]]
            cvec = torch.tensor([cw, 1.0, ai, at, as_], dtype=torch.float64)          # over (n, s, vi, vt, vs)
            cols = [tr, Rt + tr]
            # Gram matrix of the five basis vectors of this (b, k)
            G5 = torch.zeros(5, 5, dtype=torch.float64)
            tab = [tr, Rt + tr, None, None, iv]
            for i in range(5):
                for j in range(5):
                    ti, tj = tab[i], tab[j]
                    if ti is not None and tj is not None: G5[i, j] = TTm[ti, tj]
                    elif ti is not None: G5[i, j] = ST[j - 2, b, ti]
                    elif tj is not None: G5[i, j] = ST[i - 2, b, tj]
                    else: G5[i, j] = SS[(i - 2, j - 2)][b]
            s5 = torch.stack([sumT[tr], sumT[Rt + tr], sums[0, b], sums[1, b], sumT[iv]])
            su = cvec @ s5; uu = cvec @ G5 @ cvec
            mean = su / D; var = uu / D - mean * mean; rstd = (var + LN_EPS).rsqrt()
            gi = 2 if k < C else 3                                     # gg_p or gg_s
            g5 = torch.stack([ST[gi, b, tr], ST[gi, b, Rt + tr], SS[(gi, 0)][b], SS[(gi, 1)][b], ST[gi, b, iv]])   # gg . basis
            sg = sums[gi, b]
            m1 = sg / D
            m2 = rstd * (cvec @ g5 - mean * sg) / D
            # dY = al gg - be u + de 1
            al, be = rstd, rstd * rstd * m2
            de = -rstd * m1 + be * mean
            uv = G5 @ cvec                                             # u . basis
            dYv = al * g5 - be * uv + de * s5                          # dY . (n, s, vi, vt, vs)
            dYu = cvec @ dYv
            kv[b, k] = torch.stack([dYu - dYv[1], dYv[2], dYv[3], dYv[4]])            # ybar = u - s
            # ---- forward: xhat = rstd (u - mean)
            if k < C:
                cf_out[b, tr] += rstd * cw / C; cf_out[b, Rt + tr] += rstd / C
                sc_out[b] += torch.tensor([rstd * ai, rstd * at, rstd * as_, -rstd * mean], dtype=torch.float64) / C
            else:
                out_s[b] = rstd * (cw * N[tr] + S[tr] + ai * VI[b] + at * VT[b] + as_ * VSt[st] - mean)   # one row per sample: direct
            # ---- backward sums  sum_k a dY
            for q, a in enumerate((ai, at, as_)):
                cf_d[q, b, tr] += -a * be * cw; cf_d[q, b, Rt + tr] += -a * be
                sc_d[q, b] += torch.tensor([a * al if k < C else 0.0, a * al if k == C else 0.0,
                                            -a * be * ai, -a * be * at, -a * be * as_, a * de], dtype=torch.float64)
    NSt = torch.cat([N, S], 0)                                        # [2Rt, D]
    vs_b = VSt[c["sid"]]
    out_p = cf_out @ NSt + sc_out[:, 0:1] * VI + sc_out[:, 1:2] * VT + sc_out[:, 2:3] * vs_b + sc_out[:, 3:4]
    d = []
    for q in range(3):
        d.append(cf_d[q] @ NSt + sc_d[q, :, 0:1] * GGp + sc_d[q, :, 1:2] * GGs + sc_d[q, :, 2:3] * VI +
                 sc_d[q, :, 3:4] * VT + sc_d[q, :, 4:5] * vs_b + sc_d[q, :, 5:6])
    return {"out_proto": out_p * c["gamma"] + c["beta"], "out_state": out_s * c["gamma"] + c["beta"],
            "dVI": d[0], "dVT": d[1], "dVS": d[2], "kv": kv}
