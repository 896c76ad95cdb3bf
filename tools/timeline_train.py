"""Kernel timeline of ONE graph-replayed training step (train.TrainStep) from CUPTI records: python tools/timeline_train.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head, train   # noqa: E402
from oracle import synth                  # noqa: E402

B, T, C = 1024, 10, 20
capi.require_device()
dev = torch.device("cuda")
params = {k: v.to(dev) for k, v in synth.make_params(T, seed=42, perturb_ln=False).items()}
protos = synth.make_prototypes(C).to(dev)
text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
evo = [torch.randn(512, generator=torch.Generator().manual_seed(5)) for _ in range(C)]
ts = train.TrainStep(params, protos, B, text_cls, mode=head.MODE_BF16, evolution_features=evo)
b = {k: v.to(dev) for k, v in synth.make_batch(B, C, step=0).items()}
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(5):
        ts.load(b["image"], b["text"], b["state"], b["label"]); ts.step(epoch=0)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            ts.load(b["image"], b["text"], b["state"], b["label"]); ts.step(epoch=0)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
names = [e.name.split("(")[0].replace("team::", "").replace("void ", "") for e in ev]
firsts = [i for i, n in enumerate(names) if n.startswith("prep_kernel")]
lo, hi = firsts[2], firsts[3]
t0 = ev[lo].time_range.start
print(f"{hi - lo} device activities, {ev[hi].time_range.start - t0:.1f} us from prep_kernel to the next step's prep_kernel")
for i in range(lo, hi):
    e = ev[i]
    print(f"{e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:6.1f} us  {names[i][:60]}")
