"""GPU unit tests of the two GEMM engines through the C ABI: the fp32 FFMA GEMM and the
tcgen05/TMA/TMEM bf16 GEMM (all four operand-major combinations, ragged M/N/K, split-K,
alpha/beta/bias epilogue, two-term bf16 split of A).  Reference: fp64 matmul of the same
(bf16-rounded) operands, so the only difference is fp32 accumulation order -> 2e-6."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (130, 70, 33), (2048, 512, 512), (512, 512, 2048), (144, 144, 512)])
def test_gemm_f32(ta, tb, M, N, K):
    from team_b200 import capi
    capi.require_device()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tb else (K, N), generator=g).cuda()
    Cm = torch.randn((M, N), generator=g).cuda()
    bias = torch.randn((N,), generator=g).cuda()
    ref = 0.5 * ((A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())) + bias.double() + 2.0 * Cm.double()
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    capi.check(capi.lib().team_gemm_f32(ta, tb, M, N, K, 0.5, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                        2.0, Cm.data_ptr(), N, bias.data_ptr(), ws.data_ptr(), ws.numel(), _st()))
    assert rel(Cm, ref) < 2e-6


SHAPES = [(128, 64, 64), (128, 128, 128), (256, 192, 512), (2048, 1536, 512), (200, 72, 136), (144, 144, 512),
          (512, 512, 2048), (2048, 144, 512), (144, 512, 2048), (1000, 520, 200)]


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bf16_tcgen05(a_mn, b_mn, M, N, K):
    from team_b200 import capi
    capi.require_device()
    g = torch.Generator().manual_seed(7 * M + 3 * N + K + a_mn * 2 + b_mn)
    # leading dimensions must be multiples of 8 elements: pad the stored matrices
    pad = lambda n: (n + 7) // 8 * 8
    A = torch.randn((K, pad(M)) if a_mn else (M, pad(K)), generator=g).to(torch.bfloat16).cuda()
    B = torch.randn((K, pad(N)) if b_mn else (N, pad(K)), generator=g).to(torch.bfloat16).cuda()
    Av = A[:, :M] if a_mn else A[:, :K]
    Bv = B[:, :N] if b_mn else B[:, :K]
    ref = (Av.double().t() if a_mn else Av.double()) @ (Bv.double() if b_mn else Bv.double().t())
    out = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    capi.check(capi.lib().team_gemm_bf16(a_mn, b_mn, M, N, K, 1.0, A.data_ptr(), None, A.stride(0), B.data_ptr(),
                                         B.stride(0), 0.0, out.data_ptr(), N, None, ws.data_ptr(), ws.numel(), _st()),
               "team_gemm_bf16")
    torch.cuda.synchronize()
    assert rel(out, ref) < 3e-6, rel(out, ref)          # fp32 accumulation over up to 2048 products


def test_gemm_bf16_epilogue_and_split():
    from team_b200 import capi
    capi.require_device()
    M, N, K = 384, 200, 512
    g = torch.Generator().manual_seed(1)
    Af = torch.randn((M, K), generator=g).cuda()
    Bq = torch.randn((N, K), generator=g).to(torch.bfloat16).cuda()
    Cm = torch.randn((M, N), generator=g).cuda()
    bias = torch.randn((N,), generator=g).cuda()
    hi = torch.empty((M, K), dtype=torch.bfloat16, device="cuda")
    lo = torch.empty((M, K), dtype=torch.bfloat16, device="cuda")
    L = capi.lib()
    capi.check(L.team_f32_to_bf16(Af.data_ptr(), K, M, K, hi.data_ptr(), lo.data_ptr(), K, _st()))
    assert torch.equal(hi, Af.to(torch.bfloat16))
    assert torch.equal(lo, (Af - hi.float()).to(torch.bfloat16))
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    C0 = Cm.clone()
    capi.check(L.team_gemm_bf16(0, 0, M, N, K, 0.25, hi.data_ptr(), lo.data_ptr(), K, Bq.data_ptr(), K, -1.5,
                                Cm.data_ptr(), N, bias.data_ptr(), ws.data_ptr(), ws.numel(), _st()))
    ref = 0.25 * ((hi.double() + lo.double()) @ Bq.double().t()) + bias.double() - 1.5 * C0.double()
    assert rel(Cm, ref) < 2e-6
    # two-term split recovers the fp32 operand to ~2^-17
    ref32 = 0.25 * (Af.double() @ Bq.double().t()) + bias.double() - 1.5 * C0.double()
    assert rel(Cm, ref32) < 3e-5


@pytest.mark.parametrize("with_ws", [True, False])
def test_gemm_bf16_group(with_ws):
    """Grouped launch: 11 problems (two launches) with mixed operand majors, ragged shapes, long K (in-kernel
    split-K with fixed-order fold), fp32 / bf16 / dual outputs, alpha/beta/bias - each against fp64."""
    from team_b200 import capi
    capi.require_device()
    g = torch.Generator().manual_seed(11)
    pad = lambda n: (n + 7) // 8 * 8
    specs = [  # a_mn, b_mn, M, N, K, alpha, beta, bias, out ("f", "b", "fb")
        (0, 0, 2048, 144, 512, 0.5, 0.0, False, "f"), (0, 0, 2048, 144, 512, 1.0, 0.0, False, "fb"),
        (1, 1, 512, 512, 2192, 1.0, 0.0, False, "fb"), (1, 1, 144, 512, 2192, 1.0, 1.0, False, "f"),
        (0, 1, 2048, 512, 144, 1.0, 1.0, False, "fb"), (0, 0, 20, 512, 512, 1.0, 0.0, True, "f"),
        (0, 0, 10, 512, 512, 1.0, 0.0, True, "fb"), (0, 1, 512, 512, 512, 1.0, 0.0, False, "b"),
        (0, 0, 300, 72, 200, 2.0, -1.0, True, "f"), (1, 0, 130, 70, 136, 1.0, 0.0, False, "f"),
        (0, 1, 2048, 512, 1536, 1.0, 1.0, False, "f")]
    descs = (capi.GemmDesc * len(specs))()
    keep, checks = [], []
    for i, (a_mn, b_mn, M, N, K, alpha, beta, use_bias, out) in enumerate(specs):
        A = torch.randn((K, pad(M)) if a_mn else (M, pad(K)), generator=g).to(torch.bfloat16).cuda()
        B = torch.randn((K, pad(N)) if b_mn else (N, pad(K)), generator=g).to(torch.bfloat16).cuda()
        Av = A[:, :M] if a_mn else A[:, :K]
        Bv = B[:, :N] if b_mn else B[:, :K]
        C0 = torch.randn((M, N), generator=g).cuda()
        bias = torch.randn((N,), generator=g).cuda() if use_bias else None
        ref = alpha * ((Av.double().t() if a_mn else Av.double()) @ (Bv.double() if b_mn else Bv.double().t()))
        if use_bias:
            ref = ref + bias.double()
        if beta != 0.0:
            ref = ref + beta * C0.double()
        Cf = C0.clone() if "f" in out or beta != 0.0 else None
        Cb = torch.zeros((M, pad(N)), dtype=torch.bfloat16, device="cuda") if "b" in out else None
        d = descs[i]
        d.a_mn, d.b_mn, d.M, d.N, d.K, d.alpha, d.beta = a_mn, b_mn, M, N, K, alpha, beta
        d.A, d.lda, d.B, d.ldb = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
        d.C, d.ldc = (Cf.data_ptr(), N) if Cf is not None else (None, 0)
        d.C_bf16, d.ldc_bf16 = (Cb.data_ptr(), Cb.stride(0)) if Cb is not None else (None, 0)
        d.bias = bias.data_ptr() if use_bias else None
        keep += [A, B, C0, bias]
        checks.append((Cf, Cb, ref, N))
    ws = torch.empty(64 << 20 if with_ws else 1, dtype=torch.uint8, device="cuda")
    for rep in range(2):          # second round: the split-K tickets must have reset themselves
        if rep == 1:
            for i, (a_mn, b_mn, M, N, K, alpha, beta, use_bias, out) in enumerate(specs):
                if checks[i][0] is not None:
                    checks[i][0].copy_(keep[4 * i + 2])
        capi.check(capi.lib().team_gemm_bf16_group(descs, len(specs), ws.data_ptr() if with_ws else None,
                                                   ws.numel() if with_ws else 0, _st()), "team_gemm_bf16_group")
        torch.cuda.synchronize()
        for i, (Cf, Cb, ref, N) in enumerate(checks):
            if Cf is not None:
                assert rel(Cf, ref) < 3e-6, (rep, i, rel(Cf, ref))
            if Cb is not None:
                assert rel(Cb[:, :N].float(), ref) < 4e-3, (rep, i)
                assert torch.equal(Cb[:, :N], (Cf if Cf is not None else ref.float()).to(torch.bfloat16)) or Cf is None


@pytest.mark.parametrize("maj", [((1, 1), (1, 1)), ((1, 1), (0, 1)), ((0, 0), (1, 0)), ((0, 1), (0, 1))])
def test_gemm_bf16_two_segments(maj):
    """C = A1 B1 + A2 B2 accumulated in one TMEM tile (segments may differ in operand majors and in K)."""
    from team_b200 import capi
    capi.require_device()
    g = torch.Generator().manual_seed(5)
    pad = lambda n: (n + 7) // 8 * 8
    cases = [(144, 512, 2048, 144), (512, 512, 2048, 20), (512, 512, 10, 1000), (2048, 512, 1536, 72)]
    descs = (capi.GemmDesc * len(cases))()
    keep, refs, outs = [], [], []
    for i, (M, N, K1, K2) in enumerate(cases):
        ref = torch.zeros((M, N), dtype=torch.float64)
        ptrs = []
        for (a_mn, b_mn), K in zip(maj, (K1, K2)):
            A = torch.randn((K, pad(M)) if a_mn else (M, pad(K)), generator=g).to(torch.bfloat16).cuda()
            B = torch.randn((K, pad(N)) if b_mn else (N, pad(K)), generator=g).to(torch.bfloat16).cuda()
            Av = (A[:, :M] if a_mn else A[:, :K]).double().cpu()
            Bv = (B[:, :N] if b_mn else B[:, :K]).double().cpu()
            ref += (Av.t() if a_mn else Av) @ (Bv if b_mn else Bv.t())
            ptrs.append((A, B))
            keep += [A, B]
        out = torch.full((M, N), float("nan"), device="cuda")
        d = descs[i]
        (d.a_mn, d.b_mn), (d.a_mn2, d.b_mn2) = maj
        d.M, d.N, d.K, d.K2, d.alpha, d.beta = M, N, K1, K2, 1.0, 0.0
        d.A, d.lda, d.B, d.ldb = ptrs[0][0].data_ptr(), ptrs[0][0].stride(0), ptrs[0][1].data_ptr(), ptrs[0][1].stride(0)
        d.A2, d.lda2, d.B2, d.ldb2 = ptrs[1][0].data_ptr(), ptrs[1][0].stride(0), ptrs[1][1].data_ptr(), ptrs[1][1].stride(0)
        d.C, d.ldc = out.data_ptr(), N
        refs.append(ref); outs.append(out)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    capi.check(capi.lib().team_gemm_bf16_group(descs, len(cases), ws.data_ptr(), ws.numel(), _st()), "team_gemm_bf16_group")
    torch.cuda.synchronize()
    for i in range(len(cases)):
        assert rel(outs[i], refs[i]) < 3e-6, (i, rel(outs[i], refs[i]))


@pytest.mark.parametrize("two_seg", [False, True])
def test_gemm_bf16_persistent_group(two_seg):
    """Large grouped launch -> the persistent kernel (128 x 256 tiles, double-buffered accumulator, ticketed
    split-K for the long-K weight-gradient shapes): mixed majors, ragged M / N, dual outputs, alpha/beta/bias,
    two K-segments; twice, so the tickets must have reset themselves; results bit-identical between the runs."""
    from team_b200 import capi
    capi.require_device()
    g = torch.Generator().manual_seed(23)
    pad = lambda n: (n + 7) // 8 * 8
    specs = [  # a_mn, b_mn, M, N, K, alpha, beta, bias, out
        (0, 0, 8192, 1536, 512, 1.0, 0.0, False, "b"), (0, 0, 8000, 512, 512, 1.0, 0.0, True, "fb"),
        (0, 0, 8192, 144, 512, 0.5, 0.0, False, "f"), (1, 1, 512, 512, 8192 + 144, 1.0, 0.0, False, "f"),
        (1, 1, 144, 512, 8192, 1.0, 1.0, False, "fb"), (0, 1, 8192, 512, 144, 1.0, 1.0, False, "f"),
        (0, 1, 4100, 512, 1536, 1.0, 1.0, False, "f"), (1, 0, 256, 148, 4096, 2.0, 0.0, True, "f")]
    descs = (capi.GemmDesc * len(specs))()
    keep, checks = [], []
    for i, (a_mn, b_mn, M, N, K, alpha, beta, use_bias, out) in enumerate(specs):
        A = (0.25 * torch.randn((K, pad(M)) if a_mn else (M, pad(K)), generator=g)).to(torch.bfloat16).cuda()
        B = (0.25 * torch.randn((K, pad(N)) if b_mn else (N, pad(K)), generator=g)).to(torch.bfloat16).cuda()
        Av = A[:, :M] if a_mn else A[:, :K]
        Bv = B[:, :N] if b_mn else B[:, :K]
        C0 = torch.randn((M, N), generator=g).cuda()
        bias = torch.randn((N,), generator=g).cuda() if use_bias else None
        ref = alpha * ((Av.double().t() if a_mn else Av.double()) @ (Bv.double() if b_mn else Bv.double().t()))
        d = descs[i]
        if two_seg and i in (3, 6):            # second K-segment with the other operand majors
            K2 = 200
            A2 = (0.25 * torch.randn((M, pad(K2)) if a_mn else (K2, pad(M)), generator=g)).to(torch.bfloat16).cuda()
            B2 = (0.25 * torch.randn((N, pad(K2)) if b_mn else (K2, pad(N)), generator=g)).to(torch.bfloat16).cuda()
            A2v = A2[:, :K2] if a_mn else A2[:, :M]
            B2v = B2[:, :K2] if b_mn else B2[:, :N]
            ref = ref + alpha * ((A2v.double() if a_mn else A2v.double().t()) @ (B2v.double().t() if b_mn else B2v.double()))
            d.a_mn2, d.b_mn2, d.K2 = 1 - a_mn, 1 - b_mn, K2
            d.A2, d.lda2, d.B2, d.ldb2 = A2.data_ptr(), A2.stride(0), B2.data_ptr(), B2.stride(0)
            keep += [A2, B2]
        if use_bias:
            ref = ref + bias.double()
        if beta != 0.0:
            ref = ref + beta * C0.double()
        Cf = C0.clone() if "f" in out or beta != 0.0 else None
        Cb = torch.zeros((M, pad(N)), dtype=torch.bfloat16, device="cuda") if "b" in out else None
        d.a_mn, d.b_mn, d.M, d.N, d.K, d.alpha, d.beta = a_mn, b_mn, M, N, K, alpha, beta
        d.A, d.lda, d.B, d.ldb = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
        d.C, d.ldc = (Cf.data_ptr(), N) if Cf is not None else (None, 0)
        d.C_bf16, d.ldc_bf16 = (Cb.data_ptr(), Cb.stride(0)) if Cb is not None else (None, 0)
        d.bias = bias.data_ptr() if use_bias else None
        keep += [A, B, bias]
        checks.append((Cf, Cb, ref, N, C0))
    ws = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    first = None
    for rep in range(2):
        for Cf, Cb, ref, N, C0 in checks:
            if Cf is not None:
                Cf.copy_(C0)
        capi.check(capi.lib().team_gemm_bf16_group(descs, len(specs), ws.data_ptr(), ws.numel(), _st()), "team_gemm_bf16_group")
        torch.cuda.synchronize()
        for i, (Cf, Cb, ref, N, C0) in enumerate(checks):
            if Cf is not None:
                assert rel(Cf, ref) < 3e-6, (rep, i, rel(Cf, ref))
            if Cb is not None:
                assert rel(Cb[:, :N].float(), ref) < 4e-3, (rep, i)
        snap = [x.clone() for c in checks for x in c[:2] if x is not None]
        if first is None:
            first = snap
        else:
            assert all(torch.equal(a, b) for a, b in zip(first, snap))
