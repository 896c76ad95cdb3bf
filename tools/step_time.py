"""CUDA-event time of the eager head step (fwd+bwd, bf16 mode, T=10) at one batch size: python tools/step_time.py [batch] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head   # noqa: E402
from oracle import synth           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, C = 10, 20
capi.require_device()
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    runner.step(img, txt, sid, text_cls, cots)
e1.record()
torch.cuda.synchronize()
print(f"batch {B}: {e0.elapsed_time(e1) / steps:.4f} ms/step (eager, {steps} steps)  env " +
      " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("TEAM_")))
