"""Diagnostic: where does the host time of the autograd (public API) path go?"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import torch
from oracle import synth
from team_b200 import head
T, B = 10, 1024
C = 2 * T
dev = torch.device("cuda")
params = synth.make_params(T, seed=42, perturb_ln=False)
p = {k: v.to(dev).requires_grad_(v.dim() > 0) for k, v in params.items()}
pack = head.HeadParamPack.from_state_dict(p)
protos = synth.make_prototypes(C).to(dev)
b = synth.make_batch(B, C)
cots = [c.to(dev) for c in synth.make_cotangents(B)]
x, t, s, tc = b["image"].to(dev), b["text"].to(dev), b["state"].to(dev), b["text_cls"].to(dev)
mode = head.MODE_BF16 if len(sys.argv) < 2 else int(sys.argv[1])
def step():
    o = head.forward_tri_modal(pack, x, t, s, protos, text_cls=tc, mode=mode)
    torch.autograd.backward(o[:4], cots)
    return o
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
