"""team_b200.train.TrainStep - the learner's inner-loop body (models/proof.py:403-451) as one replayed CUDA graph -
against the fp64 oracle of the same loop body (reference-pinned oracle functions + torch.optim.AdamW):
loss values of every step and the parameters after three optimisation steps over two epochs (the second epoch
re-captures with the new learning rate and unicl temperature; the Adam step count lives on the device)."""
import math

import pytest
import torch

from oracle import synth
from oracle import team_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_train_step_graph_matches_the_oracle_loop():
    from team_b200 import head, train
    T, B, C = 3, 48, 6
    params = synth.make_params(T, seed=77)
    protos = synth.make_prototypes(C, seed=5)
    text_cls = synth.make_text_class_features(20)[:C].contiguous()
    g = torch.Generator().manual_seed(9)
    evo = [torch.randn(512, generator=g) if c != 2 else None for c in range(C)]
    batches = [synth.make_batch(B, C, step=s) for s in range(3)]
    epochs = [0, 0, 1]
    lr0, wd, tuned = 0.004, 0.05, 20
    ls = math.exp(2.6592600369327783)
    # ---- oracle loop (fp64, CPU): autograd + torch.optim.AdamW, exactly the learner's loop body
    p64 = {k: v.double().clone() for k, v in params.items()}
    names = O.trainable_names(params)
    for n in names:
        p64[n].requires_grad_(True)
    opt = torch.optim.AdamW([p64[n] for n in names], lr=lr0, weight_decay=wd)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=tuned, eta_min=1e-8)
    ref_losses = []
    last_epoch = 0
    for b, ep in zip(batches, epochs):
        if ep != last_epoch:
            sched.step(); last_epoch = ep
        img, txt, sid, y = b["image"].double(), b["text"].double(), b["state"], b["label"]
        with torch.no_grad():
            ce = torch.nn.functional.cross_entropy(O.forward_for_classification(p64, img, text_cls.double()), y)
        o = O.forward_tri_modal(p64, img, txt, sid, protos.double())
        un, inst, _ = O.unicl_loss(o[0], o[1], o[2], y, epoch=ep, max_epoch=tuned, state_ids=sid,
                                   evolution_features=[None if e is None else e.double() for e in evo])
        ei = torch.nn.functional.normalize(O.encode_image(img, p64, normalize=True), dim=1)
        et = torch.nn.functional.normalize(O.encode_text(txt, p64, normalize=True), dim=1)
        cl = O.clip_loss(ei, et, ls)
        total = ce + cl + 0.3 * un
        opt.zero_grad()
        total.backward()
        opt.step()
        ref_losses.append([float(total), float(ce), float(cl), float(un), float(inst)])
    # ---- the captured step
    dev = torch.device("cuda")
    pg = {k: v.clone().to(dev) for k, v in params.items()}
    ts = train.TrainStep(pg, protos.to(dev), B, text_cls.to(dev), mode=head.MODE_F32, init_lr=lr0, min_lr=1e-8,
                         weight_decay=wd, tuned_epoch=tuned, logit_scale=ls, evolution_features=evo)
    for i, (b, ep) in enumerate(zip(batches, epochs)):
        ts.load(b["image"], b["text"], b["state"], b["label"])
        got = ts.step(epoch=ep).cpu().double()
        for a, r in zip(got.tolist(), ref_losses[i]):
            assert abs(a - r) < 2e-5 * max(1.0, abs(r)), (i, got.tolist(), ref_losses[i])
    assert int(ts.opt.step_dev.cpu()) == 3
    assert len(ts._graphs) == 2                                    # one graph per epoch
    for n in names:
        moved = (p64[n].detach() - params[n].double()).norm()
        err = (pg[n].detach().double().cpu() - p64[n].detach()).norm()
        assert float(err) < 2e-3 * float(moved) + 1e-7, (n, float(err), float(moved))
