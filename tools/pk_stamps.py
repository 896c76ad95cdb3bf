"""Timeline of the persistent grouped-GEMM launches of one head step (in-kernel globaltimer stamps).
    python tools/pk_stamps.py [batch]
per CTA and item: 0 producer starts item, 1 producer issued last k-block, 2 MMA sees first k-block, 3 MMA committed
the accumulator, 4 epilogue sees the accumulator, 5 epilogue done; slot 7 = k-blocks | problem << 32."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head   # noqa: E402
from oracle import synth           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T, C = 10, 20
capi.require_device()
L = capi.lib()
L.team_gemm_debug_stamps.argtypes = [ctypes.c_void_p]
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
dbg = torch.zeros((32, 1024 * 16), dtype=torch.int64, device=dev)
L.team_gemm_debug_stamps(dbg.data_ptr())
runner.step(img, txt, sid, text_cls, cots)
torch.cuda.synchronize()
L.team_gemm_debug_stamps(None)
d = dbg.cpu()
for l in range(32):
    x = d[l][:148 * 64].view(148, 8, 8)
    if int((x[:, :, 4] != 0).sum()) == 0:
        continue
    t0 = int(x[:, 0, 0][x[:, 0, 0] != 0].min())
    tend = int(x[:, :, 5].max())
    print(f"launch {l}: span of the first 8 items {tend - t0} ns")
    for c in (0, 73, 147):
        for i in range(8):
            if x[c, i, 4] == 0:
                continue
            v = [int(x[c, i, s]) - t0 if x[c, i, s] else None for s in range(6)]
            meta = int(x[c, i, 7])
            print(f"   cta {c:3d} item {i}: prob {meta >> 32} kb {meta & 0xffffffff:4d}  prod {v[0]}..{v[1]}  mma {v[2]}..{v[3]}  epi {v[4]}..{v[5]}")
