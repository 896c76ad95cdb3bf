"""Per-launch time and TFLOP/s of every grouped-GEMM launch of one eager head step (library event records).
    python tools/gemm_waves.py [batch]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head   # noqa: E402
from oracle import synth           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = 10
C = 2 * T
capi.require_device()
L = capi.lib()
L.team_prof_dump.restype = ctypes.c_longlong
L.team_prof_dump.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong]
dev = torch.device("cuda")
params = synth.make_params(T, seed=42, perturb_ln=False)
pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
protos = synth.make_prototypes(C).to(dev)
text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
g = torch.Generator().manual_seed(B)
img = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).to(dev)
txt = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=-1).to(dev)
sid = torch.tensor([1, 3, 4])[torch.randint(0, 3, (B,), generator=g)].to(dev)
cots = [torch.randn(B, 512, generator=g).to(dev) for _ in range(4)]
runner = head.HeadStepRunner(pack, protos, B, C, head.MODE_BF16)
for _ in range(3):
    runner.step(img, txt, sid, text_cls, cots)
torch.cuda.synchronize()
L.team_prof_enable(1)
runner.step(img, txt, sid, text_cls, cots)
L.team_prof_enable(0)
ms = (ctypes.c_double * 64)(); fl = (ctypes.c_double * 64)(); kd = (ctypes.c_int * 64)()
n = L.team_prof_dump(ms, fl, kd, 64)
tot = 0.0
for i in range(n):
    tot += ms[i]
    print(f"launch {i}: kind {kd[i]}  {ms[i] * 1e3:9.1f} us  {fl[i] / 1e9:9.2f} GFLOP  {fl[i] / max(ms[i], 1e-9) / 1e9:8.1f} TFLOP/s")
print(f"batch {B}: {n} GEMM launches, {tot * 1e3:.1f} us")
