"""Golden-vector case table shared by ``oracle/gen_golden.py`` and ``tests/``
(TEST INFRASTRUCTURE).  A case is fully determined by the seeds below; the npz files
under ``tests/golden`` hold only what the REAL reference produced for it."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from oracle import synth

GRAD_ROW_STRIDE = 64     # [512,512] gradients are stored as rows 0,64,128,...

CASES: Dict[str, dict] = {
    # forward_tri_modal + autograd grads + forward_for_classification
    "head_T1_B6": {"kind": "head", "T": 1, "B": 6, "seed": 42, "step": 0},
    "head_T3_B5_5state": {"kind": "head", "T": 3, "B": 5, "seed": 43, "step": 1, "five_state": True},
    "head_T2_B7_classtext": {"kind": "head", "T": 2, "B": 7, "seed": 44, "step": 2, "class_text": True},
    "head_T10_B4": {"kind": "head", "T": 10, "B": 4, "seed": 45, "step": 3},
    # Proof_Net.forward (PROOF fusion)
    "proof_T2_B5": {"kind": "proof_forward", "T": 2, "B": 5, "seed": 46, "step": 4},
    # Proof_Net.forward WITH autograd (cotangents on image / text / proto outputs, gradients of every trainable parameter)
    "proof_T2_B5_grad": {"kind": "proof_grad", "T": 2, "B": 5, "seed": 49, "step": 5},
    # MultiHeadAttention.forward(q, k, v) standalone (cross-attention shapes), outputs + input / parameter gradients
    "mha_cross": {"kind": "mha", "B": 3, "Lq": 5, "Lk": 7, "seed": 50},
    # BaseLearner._construct_exemplar (herding) of the real reference: 3 classes x 150 rows, 12 exemplars per class
    "herding": {"kind": "herding", "n_classes": 3, "n_train": 150, "m": 12, "seed": 4321},
    # CosineLinear
    "cosine_linear": {"kind": "cosine_linear", "N": 96, "num_classes": 20, "seed": 3000, "sigma": 1.0},
    # cal_prototype / replace_fc
    "cal_prototype": {"kind": "cal_prototype", "N": 700, "num_classes": 6, "known": 2, "seed": 2001,
                      "loader_batch": 128, "empty_class": 4},
    "simplecil": {"kind": "simplecil", "N": 900, "num_classes": 20, "seed": 2002, "loader_batch": 256,
                  "zipf": True},
    # evolve_and_update + evolve_state_prototypes + sync + distance EMA
    "evolve_6cls": {"kind": "evolve", "T": 3, "num_classes": 6, "seed": 47, "proto_seed": 1006},
    "evolve_20cls": {"kind": "evolve", "T": 10, "num_classes": 20, "seed": 48, "proto_seed": 1007},
    "state_distance_forward": {"kind": "state_distance_forward", "B": 64, "seed": 4000},
    "dynamic_gcn": {"kind": "dynamic_gcn", "N": 12, "E": 30, "seed": 5000},
    # losses of the training step: unicl_loss (evolution_features=None) and ClipLoss, values + input grads
    "unicl_B24": {"kind": "unicl", "B": 24, "C": 6, "seed": 6000, "epoch": 3, "max_epoch": 10},
    "unicl_B9_static_tau": {"kind": "unicl", "B": 9, "C": 20, "seed": 6001, "epoch": None, "max_epoch": None},
    "clip_B16": {"kind": "clip", "B": 16, "seed": 6002, "logit_scale": 14.285714},
    # unicl_loss WITH evolution features (models/proof.py:51-106): 5 classes, class 3 has no feature, class ids >= 4 out of range
    "unicl_B20_evolution": {"kind": "unicl", "B": 20, "C": 6, "seed": 6003, "epoch": 5, "max_epoch": 20, "evolution": True},
    # the complete reference learner: 2 incremental tasks x 2 epochs on the synthetic DataManager (oracle/learner_harness.py)
    "learner_2tasks": {"kind": "learner", "tasks": 2, "epochs": 2, "seed": 7},
}


def grad_subsample(g: torch.Tensor) -> torch.Tensor:
    if g is None:
        return torch.zeros(0)
    if g.dim() == 2 and g.shape[0] == 512 and g.shape[1] == 512:
        return g[::GRAD_ROW_STRIDE].contiguous()
    return g


def case_inputs(case: dict) -> dict:
    kind = case["kind"]
    if kind == "herding":
        from oracle import learner_harness
        return {"data": learner_harness.FakeData(n_classes=case["n_classes"], n_train=case["n_train"], n_test=2, seed=case["seed"])}
    if kind == "mha":
        g = torch.Generator(device="cpu").manual_seed(case["seed"])
        B, Lq, Lk = case["B"], case["Lq"], case["Lk"]
        return {"params": synth.make_params(1, seed=case["seed"]),
                "q": torch.randn((B, Lq, 512), generator=g), "k": torch.randn((B, Lk, 512), generator=g),
                "v": torch.randn((B, Lk, 512), generator=g), "cot": torch.randn((B, Lq, 512), generator=g)}
    if kind in ("head", "proof_forward", "proof_grad"):
        T, B = case["T"], case["B"]
        C = synth.CLASSES_PER_TASK * T
        return {"params": synth.make_params(T, seed=case["seed"]),
                "protos": synth.make_prototypes(C, seed=1005 + case["seed"]),
                "batch": synth.make_batch(B, C, step=case["step"], five_state=case.get("five_state", False)),
                "cots": synth.make_cotangents(B, step=case["step"]), "C": C}
    if kind == "cosine_linear":
        g = torch.Generator(device="cpu").manual_seed(case["seed"])
        x, _, _ = synth.make_prototype_build_inputs(case["N"], case["num_classes"], seed=case["seed"] + 1,
                                                    normalize=False)
        bound = 1.0 / math.sqrt(512)
        w = (torch.rand((case["num_classes"], 512), generator=g) * 2 - 1) * bound
        return {"x": x, "weight": w}
    if kind in ("cal_prototype", "simplecil"):
        x, y, s = synth.make_prototype_build_inputs(
            case["N"], case["num_classes"], seed=case["seed"], normalize=(kind == "cal_prototype"),
            zipf=case.get("zipf", False), empty_class=case.get("empty_class"))
        return {"x": x, "y": y, "s": s}
    if kind == "evolve":
        C = case["num_classes"]
        return {"params": synth.make_params(case["T"], seed=case["seed"]),
                "protos": synth.make_prototypes(C, seed=1005 + case["seed"]),
                "by_state": synth.make_state_prototype_dict(C, seed=case["proto_seed"])}
    if kind == "state_distance_forward":
        g = torch.Generator(device="cpu").manual_seed(case["seed"])
        feat = torch.randn((case["B"], 512), generator=g)
        sid = torch.tensor([0, 1, 3, 4, 2, 4, 1, 4], dtype=torch.int64)[
            torch.randint(0, 8, (case["B"],), generator=g)]
        return {"feat": feat, "sid": sid}
    if kind in ("unicl", "clip"):
        g = torch.Generator(device="cpu").manual_seed(case["seed"])
        B = case["B"]
        base = torch.randn((B, 512), generator=g)
        noise = 2.0 if kind == "clip" else 0.7                                          # loss values of order 1
        feats = [base + noise * torch.randn((B, 512), generator=g) for _ in range(3)]    # correlated, un-normalised
        if kind == "clip":
            return {"image": F.normalize(feats[0], dim=1), "text": F.normalize(feats[1], dim=1)}
        labels = torch.randint(0, case["C"], (B,), generator=g, dtype=torch.int64)
        states = torch.tensor([1, 3, 4], dtype=torch.int64)[torch.randint(0, 3, (B,), generator=g)]
        out = {"image": feats[0], "text": feats[1].reshape(B, 1, 512), "state": feats[2], "labels": labels, "states": states}
        if case.get("evolution"):
            evo = [torch.randn((512,), generator=g) for _ in range(4)]
            evo[3] = None
            out["evolution"] = evo
        return out
    if kind == "dynamic_gcn":
        g = torch.Generator(device="cpu").manual_seed(case["seed"])
        N, E = case["N"], case["E"]
        x = torch.randn((N, 512), generator=g)
        ei = torch.randint(0, N, (2, E), generator=g, dtype=torch.int64)
        ew = torch.rand((E,), generator=g)
        layers = []
        for (i, o) in ((512, 256), (256, 512)):
            b = 1.0 / math.sqrt(i)
            layers.append(((torch.rand((o, i), generator=g) * 2 - 1) * b,
                           (torch.rand((o,), generator=g) * 2 - 1) * b,
                           1.0 + 0.1 * torch.randn((o,), generator=g),
                           0.1 * torch.randn((o,), generator=g)))
        return {"x": x, "edge_index": ei, "edge_weights": ew, "layers": layers}
    raise KeyError(kind)
