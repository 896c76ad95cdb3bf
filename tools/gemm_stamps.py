"""Phase timeline of one tcgen05 GEMM launch from in-kernel globaltimer stamps."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi
capi.require_device()
L = capi.lib()
L.team_gemm_debug_stamps.argtypes = [__import__("ctypes").c_void_p]
for (M, N, K, a_mn, b_mn) in [(144, 512, 512, 0, 0), (2048, 512, 512, 0, 0), (512, 512, 2048, 1, 1)]:
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    out = torch.empty((M, N), device="cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    dbg = torch.zeros((4096, 16), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        dbg.zero_()
        L.team_gemm_debug_stamps(dbg.data_ptr())
        capi.check(L.team_gemm_bf16(a_mn, b_mn, M, N, K, 1.0, A.data_ptr(), None, A.stride(0), B.data_ptr(), B.stride(0), 0.0,
                                    out.data_ptr(), N, None, ws.data_ptr(), ws.numel(), st))
        torch.cuda.synchronize()
        L.team_gemm_debug_stamps(None)
    d = dbg.cpu()
    n = int((d[:, 0] != 0).sum())
    t0 = int(d[:n, 0].min())
    print(f"M={M} N={N} K={K} ctas={n}  (ns since first CTA start; 0 start,1 prologue,2 griddep,3 producer done,4 first full,5 acc ready,6 epilogue done,7 end,8 tmem->smem,9 partial written,10 fenced,11 folded)")
    for c in list(range(min(n, 6))) + ([n - 1] if n > 6 else []):
        print("  cta", c, [int(x) - t0 if x else None for x in d[c].tolist()[:15]])
    print("  max end", int(d[:n, 7].max()) - t0)
