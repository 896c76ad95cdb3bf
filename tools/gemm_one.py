"""One tcgen05 GEMM launch (for ncu --set full captures): python tools/gemm_one.py M N K a_mn b_mn"""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi
M, N, K, a_mn, b_mn = (int(x) for x in sys.argv[1:6])
capi.require_device()
L = capi.lib()
A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
out = torch.empty((M, N), device="cuda")
ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    capi.check(L.team_gemm_bf16(a_mn, b_mn, M, N, K, 1.0, A.data_ptr(), None, A.stride(0), B.data_ptr(), B.stride(0), 0.0,
                                out.data_ptr(), N, None, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
