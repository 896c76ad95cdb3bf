"""BASELINE configs[1]: SimpleCIL prototype build (keyed segmented sum) + cosine classifier on one B200 - achieved
HBM bandwidth against the measured copy peak.  python tools/proto_bench.py [rows]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, ops   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
capi.require_device()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
out = {"rows": N, "hbm_peak_gbs": peak}
for name, dt, e in (("fp32", torch.float32, 4), ("bf16", torch.bfloat16, 2)):
    x = torch.randn((N, 512), generator=g, device=dev, dtype=torch.float32).to(dt)          # 8.6 GB / 4.3 GB >> L2
    y = torch.randint(0, 20, (N,), generator=g, device=dev)
    s = torch.randint(0, 10, (N,), generator=g, device=dev)
    W = torch.randn((20, 512), generator=g, device=dev)

    def timed(fn, iters=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms = timed(lambda: ops.keyed_sums(x, y, num_classes=20))
    by = N * (512 * e + 8)
    out[f"segsum_class_{name}"] = {"ms": ms, "gbs": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak, "bytes_per_row": 512 * e + 8}
    ms = timed(lambda: ops.keyed_sums(x, y, s, num_classes=20))
    by = N * (512 * e + 16)
    out[f"segsum_class_state_{name}"] = {"ms": ms, "gbs": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak, "bytes_per_row": 512 * e + 16}
    ms = timed(lambda: ops.cosine_logits(x, W, want_argmax=True))
    by = N * (512 * e + 4 * 20 + 8)
    out[f"cosine_logits_argmax_{name}"] = {"ms": ms, "gbs": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak, "bytes_per_row": 512 * e + 88}
    del x
    torch.cuda.empty_cache()
print(json.dumps(out))
