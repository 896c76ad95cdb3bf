"""One launch each of the configs[1] kernels at N rows (for ncu): python tools/proto_one.py [rows] [fp32|bf16]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, ops   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
capi.require_device()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((N, 512), generator=g, device=dev, dtype=torch.float32).to(dt)
y = torch.randint(0, 20, (N,), generator=g, device=dev)
s = torch.randint(0, 10, (N,), generator=g, device=dev)
W = torch.randn((20, 512), generator=g, device=dev)
for _ in range(2):
    ops.keyed_sums(x, y, num_classes=20)
    ops.keyed_sums(x, y, s, num_classes=20)
    ops.cosine_logits(x, W, want_argmax=True)
torch.cuda.synchronize()
print("ok")
