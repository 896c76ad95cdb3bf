"""GPU parity of the drop-in class surface (team_b200.inc_net.Proof_Net & friends) used the way
models/proof.py uses the reference network: build through update_prototype / update_context_prompt /
extend_task, load the reference-initialised state_dict by NAME, call forward_tri_modal /
encode_* / forward_for_classification / evolve_state_prototypes, backprop a loss, step AdamW,
deepcopy.  Checked against the reference-generated golden vectors and the CPU oracle."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import synth
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs, grad_subsample

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class FakeCLIP(nn.Module):
    """features in -> features out (the frozen towers are outside the path)"""

    def __init__(self):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)

    def encode_image(self, x, normalize=False):
        return x

    def encode_text(self, x, normalize=False):
        return x


def build_net(T, params, protos, mode="f32"):
    from team_b200 import inc_net
    dev = torch.device("cuda")
    args = {"convnet_type": "clip", "model_name": "proof", "device": [dev], "projection_type": "pure_mlp",
            "context_prompt_length_per_task": 10, "team_mode": mode}
    net = inc_net.Proof_Net(args, False, convnet=FakeCLIP().to(dev), tokenizer=lambda texts: texts)
    for t in range(T):
        net.update_prototype(2 * (t + 1)); net.update_context_prompt(); net.extend_task()
    net.to(dev)
    missing, unexpected = net.load_state_dict({k: v for k, v in params.items()}, strict=False)
    assert not unexpected, unexpected
    # state_evolution_graph IS state_embedder (alias, utils/inc_net.py:365): its keys are filled through the first name
    ok = ("state_embedder.evolution_detector", "state_evolution_graph.", "convnet")
    assert all(m.startswith(ok) for m in missing), missing
    net.img_prototypes = protos.clone().to(dev)
    net.freeze_projection_weight_new()
    net.eval()
    return net


@pytest.mark.parametrize("name", ["head_T1_B6", "head_T10_B4"])
def test_proof_net_forward_tri_modal_vs_golden(name, golden):
    case, g = CASES[name], golden(name)
    ci = case_inputs(case)
    T = case["T"]
    net = build_net(T, ci["params"], ci["protos"])
    dev = net._device
    b = ci["batch"]
    img, txt, st, pr, ls = net.forward_tri_modal(b["image"].to(dev), b["text"].to(dev), b["state"].to(dev))
    for key, o in zip(("image", "text", "state", "proto"), (img, txt, st, pr)):
        assert tuple(o.shape) == g[key].shape, key
        assert rel(o, g[key]) < 1e-5, (key, rel(o, g[key]))
    assert abs(float(ls) - float(g["logit_scale_exp"])) < 1e-4
    loss = sum((o * c.to(dev).view_as(o)).sum() for o, c in zip((img, txt, st, pr), ci["cots"]))
    loss.backward()
    sd = dict(net.named_parameters())
    for n in O.trainable_names(ci["params"]):
        assert sd[n].grad is not None, n
        assert rel(grad_subsample(sd[n].grad), g["grad:" + n]) < 2e-5, n
    # frozen parameters of older tasks receive nothing (utils/inc_net.py:494-507)
    if T > 1:
        assert sd["projs_img.0.MLP.0.weight"].grad is None and sd["context_prompts.0"].grad is None
    # the optimiser the learner builds (models/proof.py:361) steps the parameters in place
    opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-3)
    before = sd["sel_attn.w_qs.weight"].detach().clone()
    opt.step()
    assert not torch.equal(before, sd["sel_attn.w_qs.weight"].detach())
    net2 = copy.deepcopy(net)                                   # models/proof.py:297
    assert rel(net2.encode_prototpyes(True), net.encode_prototpyes(True)) == 0.0


def test_proof_net_encode_and_classification():
    T, C, B = 3, 6, 40
    params = synth.make_params(T, seed=21)
    protos = synth.make_prototypes(C, seed=4)
    net = build_net(T, params, protos)
    dev = net._device
    b = synth.make_batch(B, C, step=5)
    for norm in (False, True):
        assert rel(net.encode_image(b["image"].to(dev), norm), O.encode_image(b["image"], params, norm)) < 1e-5
        assert rel(net.encode_text(b["text"].to(dev), norm), O.encode_text(b["text"], params, norm)) < 1e-5
        assert rel(net.encode_state(b["state"].to(dev), norm), O.encode_state(b["state"], params, norm)) < 1e-5
        assert rel(net.encode_prototpyes(norm), O.encode_prototypes(protos, params, norm)) < 1e-5
    logits, amax = net.forward_for_classification(b["image"].to(dev), b["text_cls"].to(dev))
    ref = O.forward_for_classification(params, b["image"], b["text_cls"])
    assert rel(logits, ref) < 1e-5
    top2 = ref.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-5
    assert torch.equal(amax.cpu()[safe], ref.argmax(1)[safe])
    # ClipLoss-style branch (models/proof.py:428-431): autograd through encode_image / encode_text
    p = {k: v.clone().requires_grad_(v.dim() > 0) for k, v in params.items()}
    g = torch.Generator().manual_seed(3)
    ci, ct = torch.randn(B, 512, generator=g), torch.randn(B, 512, generator=g)
    ref_i, ref_t = O.encode_image(b["image"], p, True), O.encode_text(b["text"], p, False)
    ((ref_i * ci).sum() + (ref_t * ct).sum()).backward()
    net.zero_grad()
    out_i, out_t = net.encode_image(b["image"].to(dev), True), net.encode_text(b["text"].to(dev), False)
    ((out_i * ci.to(dev)).sum() + (out_t * ct.to(dev)).sum()).backward()
    sd = dict(net.named_parameters())
    for n in (f"projs_img.{T-1}.MLP.0.weight", f"projs_img.{T-1}.MLP.0.bias", f"projs_text.{T-1}.MLP.0.weight",
              f"projs_text.{T-1}.MLP.0.bias"):
        assert rel(sd[n].grad, p[n].grad) < 2e-5, n
    assert sd["projs_img.0.MLP.0.weight"].grad is None
    # Proj_Pure_MLP on its own (convs/projections.py:7-18)
    y = net.projs_img[0](b["image"].to(dev))
    W, bb = params["projs_img.0.MLP.0.weight"], params["projs_img.0.MLP.0.bias"]
    assert rel(y, b["image"] @ W.t() + bb) < 1e-5


def test_proof_net_evolve_state_prototypes_vs_golden(golden):
    name = "evolve_6cls"
    ci, g = case_inputs(CASES[name]), golden(name)
    net = build_net(CASES[name]["T"], ci["params"], ci["protos"])
    dev = net._device
    net.img_prototypes_by_state = {c: {s: v.clone().to(dev) for s, v in sd.items()} for c, sd in ci["by_state"].items()}
    res = net.state_evolution_graph.evolve_and_update(net.img_prototypes_by_state)
    keys = [(c, s) for c, sd in res["prototypes"].items() for s in sd]
    assert keys == [tuple(k) for k in g["proto_keys"].tolist()]
    assert rel(torch.stack([res["prototypes"][c][s] for c, s in keys]), g["proto_vals"]) < 1e-5
    emb = net.evolve_state_prototypes()
    assert [i for i, e in enumerate(emb) if e is not None] == g["emb_idx"].tolist()
    assert rel(net.img_prototypes, g["img_prototypes_after_sync"]) < 1e-5
    assert net.state_evolution_graph.integrate_with_state_distance(None) is True
    # TemporalStateGCN.forward with the reference's COO arguments
    from team_b200 import graph
    gr = graph.build_evolution_graph(ci["by_state"], {})
    x = torch.stack([ci["by_state"][c][s] for c in gr.class_order for s in ci["by_state"][c].keys()]).to(dev)
    ei, ew = gr.edge_list()
    out = net.state_embedder.temporal_gcn(x, torch.from_numpy(ei).to(dev), torch.from_numpy(ew).to(dev),
                                          torch.from_numpy(gr.node_time.astype(np.float32)).view(-1, 1).to(dev))
    ref = O.temporal_state_gcn(x.cpu(), torch.from_numpy(ei), torch.from_numpy(ew),
                               torch.from_numpy(gr.node_time.astype(np.float32)).view(-1, 1), ci["params"])
    assert rel(out, ref) < 1e-5


def test_state_distance_and_cosine_modules(golden):
    from team_b200 import inc_net
    dev = torch.device("cuda")
    sdm = inc_net.AdaptiveStateDistanceMatrix(num_states=10, feature_dim=512, init_with_prior=True).to(dev)
    g = golden("evolve_6cls")
    assert np.array_equal(sdm.distance_factors.detach().cpu().numpy(), g["prior_factors"])
    assert np.array_equal(sdm.get_distance_matrix().cpu().numpy(), g["prior_matrix"])
    with torch.no_grad():                                          # models/proof.py:671-675 style access
        v = sdm.distance_factors[1, 4].item()
        sdm.distance_factors[1, 4] = 0.7 * v + 0.3 * 0.5
    assert abs(sdm.get_state_distance(1, 4).item() - (0.7 * 2.0 + 0.15)) < 1e-6
    gc = golden("cosine_linear")
    case = case_inputs(CASES["cosine_linear"])
    fc = inc_net.CosineLinear(512, CASES["cosine_linear"]["num_classes"]).to(dev)
    with torch.no_grad():
        fc.weight.copy_(case["weight"].to(dev)); fc.sigma.fill_(CASES["cosine_linear"]["sigma"])
    out = fc(case["x"].to(dev))["logits"]
    assert rel(out, gc["logits"]) < 1e-5
    # SimpleCIL: update_fc + replace_fc (models/simplecil.py:31-57)
    cs = case_inputs(CASES["simplecil"]); gs = golden("simplecil")
    net = inc_net.SimpleVitNet({"device": [dev]}, True)
    net.update_fc(CASES["simplecil"]["num_classes"])
    with torch.no_grad():
        net.fc.weight.zero_()
    net.replace_fc(cs["x"].to(dev), cs["y"].to(dev))
    assert rel(net.fc.weight.data, gs["fc_weight"]) < 1e-5
