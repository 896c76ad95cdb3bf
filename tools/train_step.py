"""One COMPLETE training step of the learner's inner loop (models/proof.py:403-451) on the library, N = 1:
   classification logits (no grad) -> forward_tri_modal -> unicl_loss (value + cotangents) -> head backward with a NULL
   prototype cotangent -> ClipLoss branch (encode_image / encode_text, loss, encode backward) -> fused AdamW.
Everything after the frozen CLIP towers; features synthetic.  Prints ms/step (CUDA events, eager launches).
    python tools/train_step.py [batch] [steps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from team_b200 import capi, head, ops   # noqa: E402
from oracle import synth                # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
T, C = 10, 20
capi.require_device()
dev = torch.device("cuda")
mode = head.MODE_BF16
params = {k: v.to(dev) for k, v in synth.make_params(T, seed=42, perturb_ln=False).items()}
pack = head.HeadParamPack.from_state_dict(params)
protos = synth.make_prototypes(C).to(dev)
text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
runner = head.HeadStepRunner(pack, protos, B, C, mode)
# trainable tensors of the newest task + the shared attention block, with their gradient views in the flat buffer
gv = runner.grad_views
P = pack.T * pack.ppt
pairs = [(params[f"projs_img.{T - 1}.MLP.0.weight"], gv["w_img"].view(512, 512)), (params[f"projs_img.{T - 1}.MLP.0.bias"], gv["b_img"]),
         (params[f"projs_text.{T - 1}.MLP.0.weight"], gv["w_text"].view(512, 512)), (params[f"projs_text.{T - 1}.MLP.0.bias"], gv["b_text"]),
         (params[f"projs_state.{T - 1}.MLP.0.weight"], gv["w_state"].view(512, 512)), (params[f"projs_state.{T - 1}.MLP.0.bias"], gv["b_state"]),
         (params[f"context_prompts.{T - 1}"], gv["prompts"].view(P, 512)[(T - 1) * pack.ppt:]),
         (params["state_embedder.state_embeddings.weight"], gv["state_emb"].view(10, 512)),
         (params["sel_attn.w_qs.weight"], gv["w_q"].view(512, 512)), (params["sel_attn.w_ks.weight"], gv["w_k"].view(512, 512)),
         (params["sel_attn.w_vs.weight"], gv["w_v"].view(512, 512)), (params["sel_attn.fc.weight"], gv["w_fc"].view(512, 512)),
         (params["sel_attn.fc.bias"], gv["b_fc"]), (params["sel_attn.layer_norm.weight"], gv["ln_g"]),
         (params["sel_attn.layer_norm.bias"], gv["ln_b"])]
for p, _ in pairs:
    p.requires_grad_(True)
opt = ops.FusedAdamW([p for p, _ in pairs], lr=1e-3, weight_decay=5e-4)
batches = [synth.make_batch(B, C, step=s) for s in range(8)]
dev_b = [{k: v.to(dev) for k, v in b.items()} for b in batches]
logit_scale = float(torch.tensor(2.6592600369327783).exp())
L = capi.lib()


def train_step(b, epoch=0):
    img, txt, sid, y = b["image"], b["text"], b["state"], b["label"]
    runner.forward(img, txt, sid, text_cls)                                      # cls logits + the four feature outputs
    losses, cots = ops.unicl_loss(runner.outs[0], runner.outs[1], runner.outs[2], y, epoch=epoch, max_epoch=20,
                                  grad_scale=0.3, mode=mode)
    runner.backward(img, txt, sid, [cots[0], cots[1], cots[2], None])            # prototype output unused by the losses
    # ClipLoss branch on the projected rows (models/proof.py:428-431); its gradients add to the newest projections
    ei = head.encode_grad(pack, "image", img, normalize=True, mode=mode)
    et = head.encode_grad(pack, "text", txt, normalize=True, mode=mode)
    closs, (gi, gt) = ops.clip_loss(ei.detach(), et.detach(), logit_scale, mode=mode)
    for p, _ in pairs[:4]:
        p.grad = None
    torch.autograd.backward([ei, et], [gi, gt])
    for p, g in pairs[:4]:
        g.add_(p.grad.reshape(g.shape))
    opt.step([g for _, g in pairs])
    return losses, closs


for s in range(3):
    train_step(dev_b[s % 8])
torch.cuda.synchronize()
c0 = L.team_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(steps):
    losses, closs = train_step(dev_b[s % 8], epoch=s % 20)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"what": "complete training step of the head (logits, forward_tri_modal, unicl_loss, backward, ClipLoss branch, AdamW)",
                  "batch": B, "tasks": T, "mode": "bf16", "ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                  "library_launches_per_step": (L.team_launch_count() - c0) / steps, "eager": True,
                  "last_losses": {"unicl_total": float(losses[0]), "clip": float(closs[0])}}))
