"""Phase timeline of every grouped-GEMM launch of one head step (in-kernel globaltimer stamps)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head
from oracle import synth
import ctypes
capi.require_device()
L = capi.lib()
L.team_gemm_debug_stamps.argtypes = [ctypes.c_void_p]
T, B = 10, 1024
C = 2 * T
dev = torch.device("cuda")
params = synth.make_params(T, seed=42, perturb_ln=False)
pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
protos = synth.make_prototypes(C).to(dev)
b = synth.make_batch(B, C, step=0)
cots = [c.to(dev).reshape(B, 512) for c in synth.make_cotangents(B, step=0)]
text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
runner = head.HeadStepRunner(pack, protos, B, C, head.MODE_BF16)
img, txt, sid = b["image"].to(dev), b["text"].to(dev), b["state"].to(dev)
for _ in range(3):
    runner.step(img, txt, sid, text_cls, cots)
torch.cuda.synchronize()
dbg = torch.zeros((32, 1024, 16), dtype=torch.int64, device=dev)
L.team_gemm_debug_stamps(dbg.data_ptr())
runner.step(img, txt, sid, text_cls, cots)
torch.cuda.synchronize()
L.team_gemm_debug_stamps(None)
d = dbg.cpu()
names = "0 start,1 prologue,2 -,3 producer done,4 first full,5 acc ready,6 epilogue done,7 end,8 tmem->smem,9 cluster synced"
print(names)
for l in range(32):
    x = d[l]
    n = int((x[:, 0] != 0).sum())
    if n == 0:
        continue
    t0 = int(x[:n, 0].min())
    end = x[:n, 7]
    slow = int(end.argmax())
    dur = (x[:n, 7] - x[:n, 0]).float()
    print(f"launch {l}: ctas={n} span={int(end.max()) - t0} ns  cta dur mean={dur.mean():.0f} max={dur.max():.0f} min={dur.min():.0f}  start spread={int(x[:n,0].max())-t0}")
    for c in (slow, int(dur.argmin())):
        print("   cta", c, [int(v) - t0 if v else None for v in x[c].tolist()[:10]])
