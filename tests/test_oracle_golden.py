"""The CPU oracle (oracle/team_oracle.py) against golden vectors produced by the REAL
reference (oracle/gen_golden.py).  This is the pin that lets the GPU parity tests trust
the oracle.  Tolerances: the oracle restates the same torch ops, so fp32 results agree to
rounding (<= 2e-6 relative, mostly bit-exact); integer outputs are exact."""
import copy

import numpy as np
import pytest
import torch

from oracle import synth
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs, grad_subsample


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


HEAD_CASES = [k for k, v in CASES.items() if v["kind"] == "head"]


@pytest.mark.parametrize("name", HEAD_CASES)
def test_forward_tri_modal_and_grads(name, golden):
    case, g = CASES[name], golden(name)
    ci = case_inputs(case)
    p = {k: v.clone().requires_grad_(True) for k, v in ci["params"].items()}
    b = ci["batch"]
    text = b["text_cls"] if case.get("class_text") else b["text"]
    outs = O.forward_tri_modal(p, b["image"], text, b["state"], ci["protos"])
    for key, o in zip(("image", "text", "state", "proto", "logit_scale_exp"), outs):
        assert tuple(o.shape) == g[key].shape, key
        assert rel_err(o.detach(), g[key]) < 2e-6, key
    names = O.trainable_names(p)
    cots = list(ci["cots"])
    if case.get("class_text"):
        cots[1] = cots[1].view(cots[1].shape[0], -1)
    grads = torch.autograd.grad(outs[:4], [p[n] for n in names], grad_outputs=cots, allow_unused=True)
    for n, gr in zip(names, grads):
        ref = g["grad:" + n]
        assert rel_err(grad_subsample(gr), ref) < 5e-6, n
    with torch.no_grad():
        logits = O.forward_for_classification(p, b["image"], b["text_cls"])
    assert rel_err(logits, g["cls_logits"]) < 2e-6
    assert np.array_equal(logits.argmax(1).numpy(), g["cls_logits"].argmax(1))


def test_text_output_shape_asymmetry(golden):
    # per-sample text -> [B,1,D]; class text -> [B,D]
    assert golden("head_T1_B6")["text"].shape == (6, 1, 512)
    assert golden("head_T2_B7_classtext")["text"].shape == (7, 512)


def test_proof_forward(golden):
    case, g = CASES["proof_T2_B5"], golden("proof_T2_B5")
    ci = case_inputs(case)
    b = ci["batch"]
    with torch.no_grad():
        img, txt, ls, pr = O.forward_proof(ci["params"], b["image"], b["text_cls"], ci["protos"])
    for key, o in (("image", img), ("text", txt), ("proto", pr), ("logit_scale_exp", ls)):
        assert tuple(o.shape) == g[key].shape
        assert rel_err(o, g[key]) < 2e-6, key


def test_proof_forward_grads(golden):
    """Proof_Net.forward with autograd (real reference): outputs and the gradient of every trainable parameter."""
    from oracle.cases import grad_subsample
    case, g = CASES["proof_T2_B5_grad"], golden("proof_T2_B5_grad")
    ci = case_inputs(case)
    b = ci["batch"]
    p = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in ci["params"].items()}
    img, txt, ls, pr = O.forward_proof(p, b["image"], b["text_cls"], ci["protos"])
    cots = [ci["cots"][0], ci["cots"][2][:txt.shape[0]], ci["cots"][3][:ci["C"]]]
    names = O.trainable_names(ci["params"])
    grads = torch.autograd.grad([img, txt, pr], [p[n] for n in names], grad_outputs=cots, allow_unused=True)
    for key, o in (("image", img), ("text", txt), ("proto", pr)):
        assert rel_err(o, g[key]) < 2e-6, key
    for n, gr in zip(names, grads):
        want = g["grad:" + n]
        if gr is None:
            assert want.size == 0 or not np.any(want), n
        else:
            assert rel_err(grad_subsample(gr), want) < 1e-5, n


def test_mha_cross_attention(golden):
    """The reference MultiHeadAttention module on q != k != v (convs/projections.py:64-87): output, input and parameter grads."""
    from oracle.cases import grad_subsample
    case, g = CASES["mha_cross"], golden("mha_cross")
    ci = case_inputs(case)
    p = {k: v.clone().requires_grad_(True) for k, v in ci["params"].items() if k.startswith("sel_attn.")}
    q, k, v = (ci[n].clone().requires_grad_(True) for n in ("q", "k", "v"))
    out = O.mha(q, k, v, p)
    assert rel_err(out, g["out"]) < 2e-6
    par = [p["sel_attn." + n] for n in ("w_qs.weight", "w_ks.weight", "w_vs.weight", "fc.weight", "fc.bias", "layer_norm.weight", "layer_norm.bias")]
    grads = torch.autograd.grad(out, [q, k, v] + par, grad_outputs=ci["cot"])
    for n, gr in zip(("q", "k", "v", "w_q", "w_k", "w_v", "w_fc", "b_fc", "ln_g", "ln_b"), grads):
        assert rel_err(grad_subsample(gr), g["grad:" + n]) < 1e-5, n


def test_philox_known_answers():
    """Random123's known-answer vectors for philox4x32-10 (kat_vectors): the generator behind the library's dropout masks."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(*[np.array([w], dtype=np.uint64) for w in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want, (ctr, [hex(int(g[0])) for g in got])
    m = O.philox_keep_mask(1 << 16, 0.1, 42, 6)
    assert abs(float(m.mean()) - 0.9) < 0.006


def test_herding(golden):
    """Exemplar herding restatement vs the real BaseLearner._construct_exemplar (models/base.py:274-343): picks exact."""
    case, g = CASES["herding"], golden("herding")
    data = case_inputs(case)["data"]
    for c in range(case["n_classes"]):
        x = data.x["train"][data.y["train"] == c].numpy()
        picked, mean, _ = O.herding_select(x, case["m"])
        assert np.array_equal(picked, g["picked"][c]), c
        assert rel_err(torch.from_numpy(mean), g["class_means"][c]) < 1e-6


def test_cosine_linear(golden):
    case, g = CASES["cosine_linear"], golden("cosine_linear")
    ci = case_inputs(case)
    out = O.cosine_linear(ci["x"], ci["weight"], torch.tensor([case["sigma"]]))
    assert rel_err(out, g["logits"]) < 2e-6
    assert np.array_equal(out.argmax(1).numpy(), g["argmax"])


def test_cal_prototype(golden):
    case, g = CASES["cal_prototype"], golden("cal_prototype")
    ci = case_inputs(case)
    C = case["num_classes"]
    protos = torch.zeros(C, 512)
    by_state = {c: {} for c in range(C)}
    emb = torch.nn.functional.normalize(ci["x"], dim=-1)     # convnet.encode_image(normalize=True)
    counts, counts_cs = O.cal_prototype(emb, ci["y"], ci["s"], case["known"], C, protos, by_state)
    assert rel_err(protos, g["img_prototypes"]) < 2e-6
    assert float(protos[case["empty_class"]].abs().sum()) == 0.0       # empty class keeps zeros
    assert float(protos[:case["known"]].abs().sum()) == 0.0            # old classes untouched
    keys = [(c, s) for c, sd in by_state.items() for s in sd]
    assert keys == [tuple(k) for k in g["by_state_keys"].tolist()]     # dict order = ascending states
    vals = torch.stack([by_state[c][s] for c, s in keys])
    assert rel_err(vals, g["by_state_vals"]) < 2e-6
    for c in range(case["known"], C):
        assert counts[c] == int((ci["y"] == c).sum())
    assert sum(counts_cs.values()) == sum(counts.values())


def test_simplecil(golden):
    case, g = CASES["simplecil"], golden("simplecil")
    ci = case_inputs(case)
    W = O.simplecil_prototypes(ci["x"], ci["y"], torch.zeros(case["num_classes"], 512))
    assert rel_err(W, g["fc_weight"]) < 2e-6
    logits = O.cosine_linear(ci["x"][:64], W, torch.ones(1))
    assert rel_err(logits, g["logits64"]) < 2e-6


@pytest.mark.parametrize("name", ["evolve_6cls", "evolve_20cls"])
def test_evolve_and_update(name, golden):
    case, g = CASES[name], golden(name)
    ci = case_inputs(case)
    p = ci["params"]
    by_state = {c: {s: v.clone() for s, v in sd.items()} for c, sd in ci["by_state"].items()}
    ncls, nst, ntime, edges, w, lt = O.build_evolution_graph(by_state)
    assert np.array_equal(np.array(edges, dtype=np.int64).T, g["edge_index"])
    assert np.array_equal(np.array(w, dtype=np.float32), g["edge_weights"])
    assert np.array_equal(np.array(ntime, dtype=np.float32).reshape(-1, 1), g["time_steps"])
    lifecycle = {}
    res = O.evolve_and_update(p, by_state, lifecycle)
    keys = [(c, s) for c, sd in res["prototypes"].items() for s in sd]
    assert keys == [tuple(k) for k in g["proto_keys"].tolist()]
    vals = torch.stack([res["prototypes"][c][s] for c, s in keys])
    assert rel_err(vals, g["proto_vals"]) < 5e-6
    # aliasing quirk: the caller's inner dicts were mutated in place
    assert all(by_state[c][s] is res["prototypes"][c][s] for c, s in keys)
    emb_idx = [i for i, e in enumerate(res["embeddings"]) if e is not None]
    assert emb_idx == g["emb_idx"].tolist()
    assert rel_err(torch.stack([res["embeddings"][i] for i in emb_idx]), g["emb_vals"]) < 5e-6
    assert list(res["lifecycle_features"].keys()) == g["lifecycle_idx"].tolist()
    assert rel_err(torch.stack(list(res["lifecycle_features"].values())), g["lifecycle_vals"]) < 2e-6
    dk = [(s1, s2) for s1, dd in res["distances"].items() for s2 in dd]
    assert dk == [tuple(k) for k in g["dist_keys"].tolist()]
    dv = np.array([res["distances"][a][b] for a, b in dk])
    assert np.allclose(dv, g["dist_vals"], rtol=0, atol=5e-6)
    # second GCN pass + no-op blend + sync
    protos = ci["protos"].clone()
    O.evolve_state_prototypes(p, protos, by_state, lifecycle)
    assert rel_err(protos, g["img_prototypes_after_sync"]) < 5e-6
    # prior constants and the Learner's double EMA
    f = O.prior_distance_factors()
    assert np.array_equal(f.numpy(), g["prior_factors"])
    assert np.array_equal(O.get_distance_matrix(f).numpy(), g["prior_matrix"])
    m = g["prior_matrix"]
    assert m[1, 4] == 2.0 and m[3, 4] == np.float32(0.7) and m[1, 2] == 1.5 and m[0, 3] == np.float32(1.8) and m[0, 0] == 1.0
    res3 = O.evolve_and_update(p, by_state, lifecycle)
    O.update_state_distance_matrix(f, res3["distances"])
    assert np.allclose(f.numpy(), g["factors_after_update"], rtol=0, atol=5e-6)
    assert np.allclose(O.get_distance_matrix(f).numpy(), g["matrix_after_update"], rtol=0, atol=5e-6)


def test_state_distance_forward(golden):
    case, g = CASES["state_distance_forward"], golden("state_distance_forward")
    ci = case_inputs(case)
    f = O.prior_distance_factors()
    ret0, cnt = O.state_distance_forward(f, ci["feat"], ci["sid"], 0)
    assert np.array_equal(ret0.numpy(), g["ret0"])        # pre-update matrix is returned
    assert np.allclose(f.numpy(), g["factors1"], rtol=0, atol=2e-6)
    ret1, cnt = O.state_distance_forward(f, ci["feat"], ci["sid"], cnt)
    assert np.allclose(ret1.numpy(), g["ret1"], rtol=0, atol=2e-6)
    assert np.allclose(f.numpy(), g["factors2"], rtol=0, atol=2e-6)
    assert cnt == int(g["counter"])
    assert f[0, 1] == np.float32(1.8)                     # state 0 never gets a centre


def test_dynamic_gcn(golden):
    case, g = CASES["dynamic_gcn"], golden("dynamic_gcn")
    ci = case_inputs(case)
    out = O.dynamic_gcn(ci["x"], ci["edge_index"], ci["edge_weights"], ci["layers"])
    assert rel_err(out, g["out"]) < 2e-6


def test_known_answer_constants():
    assert abs(512 ** 0.5 - 22.627416997969522) < 1e-12          # attention temperature
    assert O.LN_EPS == 1e-5 and O.NORM_EPS == 1e-12 and O.COS_EPS == 1e-8
    assert O.detect_evolution_type([1, 4]) == "larvae_to_adult"
    assert O.detect_evolution_type([3, 4]) == "nymph_to_adult"
    assert O.detect_evolution_type([4]) == "adult_only"
    assert O.detect_evolution_type([2, 5]) == "unknown"
    # SURVEY App. A-7 probed example: 3 classes {1,4},{3,4},{1,2,4} -> 7 nodes, 9 edges
    z = torch.zeros(512)
    bs = {0: {1: z, 4: z}, 1: {3: z, 4: z}, 2: {1: z, 2: z, 4: z}}
    _, _, _, edges, w, _ = O.build_evolution_graph(bs)
    assert list(zip([e[0] for e in edges], [e[1] for e in edges], w)) == [
        (0, 1, 0.0), (2, 3, 0.0), (4, 5, .5), (4, 6, 0.0), (5, 6, .5), (0, 4, .5), (1, 6, .5),
        (4, 0, .5), (6, 1, .5)]


@pytest.mark.parametrize("name", ["unicl_B24", "unicl_B9_static_tau"])
def test_unicl_loss(name, golden):
    """unicl_loss (models/proof.py:21-191, evolution_features=None): values and input gradients of the REAL reference."""
    case, g = CASES[name], golden(name)
    ci = case_inputs(case)
    x = [ci[k].clone().double().requires_grad_(True) for k in ("image", "text", "state")]
    total, inst, cat = O.unicl_loss(x[0], x[1], x[2], ci["labels"], epoch=case["epoch"], max_epoch=case["max_epoch"])
    total.backward()
    assert rel_err(total.detach(), g["total"]) < 1e-6
    assert abs(float(inst) - float(g["instance"])) < 1e-6 and abs(float(cat) - float(g["category"])) < 1e-6
    for k, t in zip(("g_image", "g_text", "g_state"), x):
        assert rel_err(t.grad.reshape(-1, 512), g[k]) < 2e-6, k


def test_clip_loss(golden):
    case, g = CASES["clip_B16"], golden("clip_B16")
    ci = case_inputs(case)
    x = [ci[k].clone().double().requires_grad_(True) for k in ("image", "text")]
    loss = O.clip_loss(x[0], x[1], case["logit_scale"])
    loss.backward()
    assert rel_err(loss.detach(), g["loss"]) < 1e-6
    assert rel_err(x[0].grad, g["g_image"]) < 2e-6 and rel_err(x[1].grad, g["g_text"]) < 2e-6


def test_unicl_loss_with_evolution_features(golden):
    """The evolution_features branch (models/proof.py:51-106; still CPU-only): oracle vs the real reference,
    value and input gradients (the enhancement is differentiable w.r.t. the state rows)."""
    case, g = CASES["unicl_B20_evolution"], golden("unicl_B20_evolution")
    ci = case_inputs(case)
    x = [ci[k].clone().double().requires_grad_(True) for k in ("image", "text", "state")]
    evo = [None if e is None else e.double() for e in ci["evolution"]]
    total, inst, cat = O.unicl_loss(x[0], x[1], x[2], ci["labels"], epoch=case["epoch"], max_epoch=case["max_epoch"],
                                    state_ids=ci["states"], evolution_features=evo)
    total.backward()
    assert rel_err(total.detach(), g["total"]) < 1e-6
    for k, t in zip(("g_image", "g_text", "g_state"), x):
        assert rel_err(t.grad.reshape(-1, 512), g[k]) < 2e-6, k
