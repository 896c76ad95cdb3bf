"""Seeded synthetic parameters and inputs for the TEAM head (TEST INFRASTRUCTURE).

This module is part of ``oracle/``: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py`` (input generation + the ``cpu_baseline`` / ``--impl reference``
legs) may import it.  It never touches ``/root/reference``.

Parameter names and shapes follow the reference ``state_dict`` of ``Proof_Net``
after T tasks (reference ``utils/inc_net.py:342-369``,
``convs/projections.py:44-62``, ``models/state_evolution.py:9-43``,
``models/dynamic_modal_graph.py:210-292``).  Distributions mirror the reference
initialisers (nn.Linear default bound 1/sqrt(fan_in); attention N(0, sqrt(2/(d+d)));
xavier-normal fc; N(0,1) prompts and embedding) but are drawn from *our own*
``torch.Generator`` so that both the golden-vector generator (which loads them into
the real reference modules) and the tests can rebuild identical tensors from a seed.

Input recipe follows SURVEY.md section 8(d).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

D = 512            # CLIP ViT-B/16 embed dim (reference utils/inc_net.py:21)
HID = 256          # InsectLifecycleModel hidden_dim = feature_dim // 2 (utils/inc_net.py:360)
TDIM = 64          # time encoder width = hidden_dim // 4 (models/dynamic_modal_graph.py:228)
GH = HID + TDIM    # 320, width of a TemporalGCNBlock
NUM_STATES = 10    # utils/inc_net.py:361
PROMPTS_PER_TASK = 10   # exps/IIMinsects202.json:26
CLASSES_PER_TASK = 2    # exps/IIMinsects202.json:13-14
LOGIT_SCALE_INIT = math.log(1.0 / 0.07)


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def _normal(g, shape, std=1.0):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _linear(g, out_f, in_f, prefix, out: Dict[str, torch.Tensor], bias=True):
    bound = 1.0 / math.sqrt(in_f)
    out[prefix + ".weight"] = _uniform(g, (out_f, in_f), bound)
    if bias:
        out[prefix + ".bias"] = _uniform(g, (out_f,), bound)


def _layernorm(g, n, prefix, out, perturb):
    if perturb:
        out[prefix + ".weight"] = 1.0 + 0.1 * _normal(g, (n,))
        out[prefix + ".bias"] = 0.1 * _normal(g, (n,))
    else:
        out[prefix + ".weight"] = torch.ones(n)
        out[prefix + ".bias"] = torch.zeros(n)


def make_params(T: int, seed: int = 42, perturb_ln: bool = True,
                prompts_per_task: int = PROMPTS_PER_TASK) -> Dict[str, torch.Tensor]:
    """Head parameters after T incremental tasks, keyed by reference state_dict names.

    ``perturb_ln=True`` draws non-trivial LayerNorm gains/offsets so that parity
    tests exercise gamma/beta; ``False`` reproduces the reference init (1, 0).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for kind in ("img", "text", "state"):
        for t in range(T):
            _linear(g, D, D, f"projs_{kind}.{t}.MLP.0", p)
    std_qk = math.sqrt(2.0 / (D + D))
    p["sel_attn.w_qs.weight"] = _normal(g, (D, D), std_qk)
    p["sel_attn.w_ks.weight"] = _normal(g, (D, D), std_qk)
    p["sel_attn.w_vs.weight"] = _normal(g, (D, D), std_qk)
    _layernorm(g, D, "sel_attn.layer_norm", p, perturb_ln)
    p["sel_attn.fc.weight"] = _normal(g, (D, D), math.sqrt(2.0 / (D + D)))
    p["sel_attn.fc.bias"] = _uniform(g, (D,), 1.0 / math.sqrt(D))
    for t in range(T):
        p[f"context_prompts.{t}"] = _normal(g, (prompts_per_task, D))
    p["state_embedder.state_embeddings.weight"] = _normal(g, (NUM_STATES, D))
    tg = "state_embedder.temporal_gcn"
    _linear(g, HID, D, f"{tg}.node_encoder.0", p)
    _layernorm(g, HID, f"{tg}.node_encoder.1", p, perturb_ln)
    _linear(g, TDIM, 1, f"{tg}.time_encoder.0", p)
    _layernorm(g, TDIM, f"{tg}.time_encoder.1", p, perturb_ln)
    for blk in range(2):
        b = f"{tg}.temporal_blocks.{blk}"
        _linear(g, GH, 2 * GH, f"{b}.message_net.0", p)
        _layernorm(g, GH, f"{b}.message_net.1", p, perturb_ln)
        _linear(g, GH, 2 * GH, f"{b}.update_net.0", p)
        _layernorm(g, GH, f"{b}.update_net.1", p, perturb_ln)
        _linear(g, 1, GH, f"{b}.temporal_gate.0", p)
    _linear(g, D, GH, f"{tg}.output_proj", p)
    p["convnet.logit_scale"] = torch.tensor(LOGIT_SCALE_INIT, dtype=torch.float32)
    return p


def make_class_means(num_classes: int = 20, seed: int = 1000) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return _normal(g, (num_classes, D))


def make_text_class_features(num_classes: int = 20, seed: int = 1003) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return F.normalize(_normal(g, (num_classes, D)), dim=-1)


def make_prototypes(num_classes: int, seed: int = 1005) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return F.normalize(_normal(g, (num_classes, D)), dim=-1)


def make_batch(B: int, C: int, step: int = 0, five_state: bool = False,
               normalize: bool = True, num_classes_total: int = 20,
               noise: float = 0.5) -> Dict[str, torch.Tensor]:
    """One batch of head inputs: class-clustered unit-norm image
    features, per-sample text features (= class text feature of the label), state
    ids drawn from {1,3,4} (or {1,2,3,4,5}), labels uniform over the C seen classes."""
    mu = make_class_means(num_classes_total)
    txt_cls = make_text_class_features(num_classes_total)
    gy = torch.Generator(device="cpu").manual_seed(1002 + 7919 * step)
    y = torch.randint(0, C, (B,), generator=gy, dtype=torch.int64)
    ge = torch.Generator(device="cpu").manual_seed(1001 + 7919 * step)
    x = mu[y] + noise * _normal(ge, (B, D))
    if normalize:
        x = F.normalize(x, dim=-1)
    gs = torch.Generator(device="cpu").manual_seed(1004 + 7919 * step)
    if five_state:
        states = torch.tensor([1, 3, 4, 2, 5], dtype=torch.int64)
        probs = torch.tensor([0.30, 0.15, 0.45, 0.05, 0.05])
    else:
        states = torch.tensor([1, 3, 4], dtype=torch.int64)
        probs = torch.tensor([0.35, 0.15, 0.50])
    sid = states[torch.multinomial(probs, B, replacement=True, generator=gs)]
    return {"image": x.contiguous(), "text": txt_cls[y].contiguous(), "label": y,
            "state": sid.contiguous(), "text_cls": txt_cls[:C].contiguous()}


def make_cotangents(B: int, step: int = 0) -> Tuple[torch.Tensor, ...]:
    """N(0,1) cotangents for (image, text[B,1,D], state, proto) (seeds 1010-1013)."""
    out = []
    for k, shape in enumerate(((B, D), (B, 1, D), (B, D), (B, D))):
        g = torch.Generator(device="cpu").manual_seed(1010 + k + 7919 * step)
        out.append(_normal(g, shape))
    return tuple(out)


def make_prototype_build_inputs(N: int, num_classes: int = 20, seed: int = 2000,
                                normalize: bool = True, zipf: bool = False,
                                empty_class: int | None = None):
    """Features/labels/states for the class-prototype build (a11/a12)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    mu = make_class_means(num_classes)
    if zipf:
        w = 1.0 / torch.arange(1, num_classes + 1, dtype=torch.float32)
    else:
        w = torch.ones(num_classes)
    if empty_class is not None:
        w[empty_class] = 0.0
    y = torch.multinomial(w / w.sum(), N, replacement=True, generator=g).to(torch.int64)
    x = mu[y] + 0.5 * _normal(g, (N, D))
    if normalize:
        x = F.normalize(x, dim=-1)
    states = torch.tensor([1, 3, 4], dtype=torch.int64)
    sid = states[torch.multinomial(torch.tensor([0.35, 0.15, 0.5]), N, replacement=True, generator=g)]
    return x.contiguous(), y.contiguous(), sid.contiguous()


def make_state_prototype_dict(num_classes: int = 20, seed: int = 1006,
                              pattern: Tuple[Tuple[int, ...], ...] = ((1, 4), (3, 4), (1, 2, 4)),
                              ) -> Dict[int, Dict[int, torch.Tensor]]:
    """{class: {state: unit-norm proto}} with the cyclic state pattern of SURVEY 8(d)
    (20 classes -> 46 nodes)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out: Dict[int, Dict[int, torch.Tensor]] = {}
    for c in range(num_classes):
        out[c] = {}
        for s in pattern[c % len(pattern)]:
            out[c][int(s)] = F.normalize(_normal(g, (D,)), dim=0)
    return out
