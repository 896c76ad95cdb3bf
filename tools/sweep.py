"""Batch sweep of the head fwd+bwd step on one B200 (BASELINE configs[4]: batch 256 ... 65 536).

    python tools/sweep.py [--batches 256,1024,...] [--steps 20] [--out gpurun_out/sweep.jsonl]

Per batch: CUDA-graph replay over rotating input sets (larger than L2 where they fit the slot
budget), CUDA-event timing, the survey's algorithmic flops (55.07 MF/sample + 0.47 GF/step shared
rows at T=10) against the measured sustained bf16 peak, and the grouped-GEMM share from the
library's per-launch events (eager steps after the timed region).
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head          # noqa: E402
from oracle import synth                  # noqa: E402  (input generation only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="256,512,1024,2048,4096,8192,16384,32768,65536")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tasks", type=int, default=10)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    capi.require_device()
    L = capi.lib()
    dev = torch.device("cuda")
    T = a.tasks
    C = 2 * T
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    params = synth.make_params(T, seed=42, perturb_ln=False)
    pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
    protos = synth.make_prototypes(C).to(dev)
    text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
    lines = []
    for B in [int(x) for x in a.batches.split(",")]:
        rot = max(2, min(32, (512 << 20) // (B * 512 * 4 * 6)))
        gen = torch.Generator(device="cpu").manual_seed(B)
        sets = []
        for i in range(rot):
            img = torch.nn.functional.normalize(torch.randn(B, 512, generator=gen), dim=-1).to(dev)
            txt = torch.nn.functional.normalize(torch.randn(B, 512, generator=gen), dim=-1).to(dev)
            sid = torch.tensor([1, 3, 4])[torch.randint(0, 3, (B,), generator=gen)].to(dev)
            cots = [torch.randn(B, 512, generator=gen).to(dev) for _ in range(4)]
            sets.append((img, txt, sid, cots))
        runner = head.HeadStepRunner(pack, protos, B, C, head.MODE_BF16)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for i in range(2):
                s = sets[i % rot]
                runner.step(s[0], s[1], s[2], text_cls, s[3])
            torch.cuda.synchronize()
            graphs = []
            for s in sets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    runner.step(s[0], s[1], s[2], text_cls, s[3])
                graphs.append(g)
            for i in range(5):
                graphs[i % rot].replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(a.steps):
                graphs[i % rot].replay()
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            tms, tfl, tby, nl = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
            L.team_prof_enable(1)
            npf = 3
            for i in range(npf):
                s = sets[i % rot]
                runner.step(s[0], s[1], s[2], text_cls, s[3])
            L.team_prof_enable(0)
            capi.check(L.team_prof_collect(1, ctypes.byref(tms), ctypes.byref(tfl), ctypes.byref(tby), ctypes.byref(nl)), "prof")
        falg = 55.07e6 * B + 0.47e9
        line = {"batch": B, "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "alg_tflops": falg / ms / 1e9,
                "frac_of_sustained_bf16_peak": falg / ms / 1e9 / peak, "rot_sets": rot,
                "gemm_ms_per_step_eager": tms.value / npf, "gemm_tflops": tfl.value / max(tms.value, 1e-9) / 1e9,
                "gemm_share": tms.value / npf / ms, "workspace_MiB": runner.nbytes / 2**20}
        print(json.dumps(line), flush=True)
        lines.append(line)
        del graphs, runner, sets
        torch.cuda.empty_cache()
    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
