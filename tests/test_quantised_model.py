"""CPU pins of the two fp64 models the bf16-mode GPU tests compare against:
  * oracle/quantised_model.py with every rounding switched off is the reference-pinned oracle
    (oracle/team_oracle.py, itself checked against the reference's golden vectors) - outputs AND the
    hand-written backward;
  * with ``operands_only=True`` it is that oracle evaluated on bf16-rounded inputs / weights."""
import pytest
import torch

import factorised_model as F
from oracle import quantised_model as Q
from oracle import synth
from oracle import team_oracle as O


def rel(a, b):
    a = a.detach().double(); b = b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _case(T, B, five=False):
    C = 2 * T
    params = synth.make_params(T, seed=100 + T)
    protos = synth.make_prototypes(C, seed=7)
    batch = synth.make_batch(B, C, step=T, five_state=five)
    cots = synth.make_cotangents(B, step=T)
    return params, protos, batch, cots


@pytest.mark.parametrize("T,B,five", [(1, 7, False), (3, 12, True)])
def test_models_without_rounding_equal_the_oracle(T, B, five, monkeypatch):
    params, protos, batch, cots = _case(T, B, five)
    p64 = {k: v.double().requires_grad_(v.dim() > 0) for k, v in params.items()}
    ref = O.forward_tri_modal(p64, batch["image"].double(), batch["text"].double(), batch["state"], protos.double())
    names = O.trainable_names(params)
    gref = dict(zip(names, torch.autograd.grad(ref[:4], [p64[n] for n in names], grad_outputs=[c.double() for c in cots])))
    pd = {k: v.detach() for k, v in p64.items()}
    args = (pd, batch["image"].double(), batch["text"].double(), batch["state"], protos.double(), [c.double() for c in cots])
    monkeypatch.setattr(Q, "q_bf16", lambda x: x)
    for mod, kw in ((F, {}), (Q, {"act": None, "grad": None})):
        outs, grads = mod.head_fwd_bwd(*args, **kw)
        for o, r in zip(outs, ref[:4]):
            assert rel(o.reshape(r.shape), r) < 1e-10
        for n in names:
            assert rel(grads[n], gref[n]) < 1e-9, (mod.__name__, n)


def test_operands_only_is_the_oracle_on_quantised_operands():
    T, B = 2, 9
    params, protos, batch, cots = _case(T, B)
    q = {}
    for kind in ("img", "text", "state"):                  # the projections are summed in fp32, then rounded once
        W = sum(params[f"projs_{kind}.{t}.MLP.0.weight"] for t in range(T))
        b = sum(params[f"projs_{kind}.{t}.MLP.0.bias"] for t in range(T))
        q[f"projs_{kind}.0.MLP.0.weight"] = Q.q_bf16(W).double()
        q[f"projs_{kind}.0.MLP.0.bias"] = b.double()
    q["context_prompts.0"] = torch.cat([params[f"context_prompts.{t}"] for t in range(T)], 0).double()
    for k in ("sel_attn.w_qs.weight", "sel_attn.w_ks.weight", "sel_attn.w_vs.weight", "sel_attn.fc.weight",
              "state_embedder.state_embeddings.weight"):
        q[k] = Q.q_bf16(params[k]).double()
    for k in ("sel_attn.fc.bias", "sel_attn.layer_norm.weight", "sel_attn.layer_norm.bias"):
        q[k] = params[k].double()
    q["convnet.logit_scale"] = params["convnet.logit_scale"].double()
    qb = lambda t: Q.q_bf16(t).double()
    ref = O.forward_tri_modal(q, qb(batch["image"]), qb(batch["text"]), batch["state"], qb(protos))
    lref = O.forward_for_classification(q, qb(batch["image"]), qb(batch["text_cls"]))
    p64 = {k: v.double() for k, v in params.items()}
    outs, _, logits = Q.head_fwd_bwd(p64, batch["image"].double(), batch["text"].double(), batch["state"], protos.double(),
                                     [c.double() for c in cots], text_cls=batch["text_cls"].double(), operands_only=True)
    for o, r in zip(outs, ref[:4]):                       # 1e-7: the bias sums above are fp32 sums, the model's fp64
        assert rel(o.reshape(r.shape), r) < 1e-7
    assert rel(logits, lref) < 1e-7
