"""Kernel timeline of ONE graph-replayed head step (fwd + bwd, bf16, T = 10) from CUPTI activity records (torch.profiler):
start offset, duration, stream of every kernel - shows which kernels really overlap and where the stream idles.
    python tools/timeline.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head   # noqa: E402
from oracle import synth           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T, C = 10, 20
capi.require_device()
dev = torch.device("cuda")
params = synth.make_params(T, seed=42, perturb_ln=False)
pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
protos = synth.make_prototypes(C).to(dev)
text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
sets = []
for i in range(16):
    b = synth.make_batch(B, C, step=i)
    c = synth.make_cotangents(B, step=i)
    sets.append((b["image"].to(dev), b["text"].to(dev), b["state"].to(dev), [c[0].to(dev), c[1].reshape(B, 512).to(dev), c[2].to(dev), c[3].to(dev)]))
runner = head.HeadStepRunner(pack, protos, B, C, head.MODE_BF16)
st = torch.cuda.Stream()
graphs = []
with torch.cuda.stream(st):
    runner.step(*sets[0][:3], text_cls, sets[0][3])
    torch.cuda.synchronize()
    for q in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            runner.step(q[0], q[1], q[2], text_cls, q[3])
        graphs.append(g)
    for i in range(32):
        graphs[i % 16].replay()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for i in range(6):
            graphs[i % 16].replay()
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
ev.sort(key=lambda e: e.time_range.start)
names = [e.name.split("(")[0].replace("team::", "").replace("void ", "") for e in ev]
# one step = from a prep_kernel to the next
starts = [i for i, n in enumerate(names) if n.startswith("fill_prompt_rows") or n.startswith("prep_kernel")]
firsts = [i for k, i in enumerate(starts) if k == 0 or i - starts[k - 1] > 2]
lo, hi = firsts[3], firsts[4]
t0 = ev[lo].time_range.start
print(f"step of {hi - lo} kernels, {ev[hi].time_range.start - t0:.1f} us from its first kernel to the next step's first kernel")
end_prev = t0
for i in range(lo, hi):
    e = ev[i]
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print(f"{s:8.1f} +{d:6.1f} us  stream {getattr(e, 'device_index', 0)}/{getattr(e, 'device_resource_id', '?'):>3}  {names[i][:40]}")
