"""Prints bf16-mode error statistics of the head against the fp64 oracle (diagnostic, run by hand)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import torch
from oracle import synth, team_oracle as O
from team_b200 import head

def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

def quant(params):
    q = {}
    for k, v in params.items():
        q[k] = v.to(torch.bfloat16).float() if (v.dim() == 2 and "layer_norm" not in k) else v.clone()
    return q

for (T, B) in [(1, 64), (10, 256)]:
    C = 2 * T
    params = synth.make_params(T, seed=100 + T)
    protos = synth.make_prototypes(C, seed=7)
    batch = synth.make_batch(B, C, step=T)
    cots = synth.make_cotangents(B, step=T)
    dev = torch.device("cuda")
    p = {k: v.to(dev).requires_grad_(v.dim() > 0) for k, v in params.items()}
    pack = head.HeadParamPack.from_state_dict(p)
    names = O.trainable_names(params)
    p64 = {k: v.double().requires_grad_(v.dim() > 0) for k, v in params.items()}
    ref = O.forward_tri_modal(p64, batch["image"].double(), batch["text"].double(), batch["state"], protos.double())
    gref = torch.autograd.grad(ref[:4], [p64[n] for n in names], grad_outputs=[c.double() for c in cots])
    lref = O.forward_for_classification({k: v.detach() for k, v in p64.items()}, batch["image"].double(), batch["text_cls"].double())
    for mode, nm in ((head.MODE_F32, "f32"), (head.MODE_BF16, "bf16")):
        outs = head.forward_tri_modal(pack, batch["image"].to(dev), batch["text"].to(dev), batch["state"].to(dev), protos.to(dev),
                                      text_cls=batch["text_cls"].to(dev), mode=mode)
        grads = torch.autograd.grad(outs[:4], [p[n] for n in names], grad_outputs=[c.to(dev) for c in cots])
        print(f"T={T} B={B} mode={nm}: outs", " ".join(f"{rel(a, b):.2e}" for a, b in zip(outs[:4], ref[:4])),
              "logits", f"{rel(outs[4], lref):.2e}", "argmax_match", float((outs[5].cpu() == lref.argmax(1)).float().mean()))
        print("   grads", " ".join(f"{n.split('.')[-2][:6]}.{n.split('.')[-1][:1]}={rel(g, r):.1e}" for n, g, r in zip(names, grads, gref)))
