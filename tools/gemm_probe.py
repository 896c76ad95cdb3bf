"""Microbenchmark of the tcgen05 GEMM for the shapes the head uses: N back-to-back launches captured in one
CUDA graph, timed with CUDA events (steady-state cost per launch inside a graph)."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi

def main():
    capi.require_device()
    L = capi.lib()
    st = torch.cuda.Stream()
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    shapes = [(0, 0, 2048, 512, 512), (0, 0, 2048, 1536, 512), (0, 0, 2048, 2048, 512), (1, 1, 512, 512, 2048), (0, 0, 2048, 144, 512),
              (0, 1, 2048, 512, 144), (1, 1, 144, 512, 2048), (0, 0, 144, 512, 512), (0, 0, 8192, 512, 512), (0, 0, 16384, 1536, 512)]
    reps = int(os.environ.get('REPS', '50'))
    use_ws = os.environ.get('NO_WS', '0') == '0'
    res = []
    with torch.cuda.stream(st):
        for (a_mn, b_mn, M, N, K) in shapes:
            A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
            B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
            outs = [torch.empty((M, N), device="cuda") for _ in range(4)]
            def launch(i):
                capi.check(L.team_gemm_bf16(a_mn, b_mn, M, N, K, 1.0, A.data_ptr(), None, A.stride(0), B.data_ptr(), B.stride(0),
                                            0.0, outs[i % 4].data_ptr(), N, None, ws.data_ptr() if use_ws else None, ws.numel() if use_ws else 0, st.cuda_stream))
            launch(0); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for i in range(reps):
                    launch(i)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); g.replay(); e1.record(st); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (2 * reps)
            res.append({"a_mn": a_mn, "b_mn": b_mn, "M": M, "N": N, "K": K, "us_per_launch": round(us, 2),
                        "tflops": round(2.0 * M * N * K / us / 1e6, 1)})
            print(res[-1], flush=True)
        # conversion kernel
        X = torch.randn(2048, 512, device="cuda"); Xb = torch.empty(2048, 512, device="cuda", dtype=torch.bfloat16)
        g = torch.cuda.CUDAGraph()
        L.team_f32_to_bf16(X.data_ptr(), 512, 2048, 512, Xb.data_ptr(), None, 512, st.cuda_stream); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(reps):
                L.team_f32_to_bf16(X.data_ptr(), 512, 2048, 512, Xb.data_ptr(), None, 512, st.cuda_stream)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
        print({"f32_to_bf16_2048x512_us": round(e0.elapsed_time(e1) * 1e3 / reps, 2)})
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w"))

if __name__ == "__main__":
    main()
