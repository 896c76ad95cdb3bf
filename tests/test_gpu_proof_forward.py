"""GPU parity of the PROOF fusion forward (Proof_Net.forward / forward_transformer, utils/inc_net.py:436-492;
SURVEY 8a row a8) through team_head_proof_fwd against the reference-generated golden vector and the fp64 oracle.
fp32 mode bar: <= 1e-5 relative (norm-wise); bf16 mode: <= 1e-2 against the unquantised fp64 oracle."""
import pytest
import torch

from oracle import synth
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _pack(params, dev):
    from team_b200 import head
    return head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})


def test_proof_forward_vs_golden(golden):
    from team_b200 import head
    dev = torch.device("cuda")
    case, g = CASES["proof_T2_B5"], golden("proof_T2_B5")
    ci = case_inputs(case)
    b = ci["batch"]
    img, txt, pro = head.forward_proof(_pack(ci["params"], dev), b["image"].to(dev), b["text_cls"].to(dev),
                                       ci["protos"].to(dev), mode=head.MODE_F32)
    for key, o in (("image", img), ("text", txt), ("proto", pro)):
        assert tuple(o.shape) == g[key].shape, key
        assert rel(o, g[key]) < 1e-5, (key, rel(o, g[key]))


@pytest.mark.parametrize("T,B,Tn", [(1, 1, 2), (3, 37, 6), (10, 300, 20), (4, 1029, 5)])
def test_proof_forward_vs_oracle(T, B, Tn):
    """Ragged batches (1, 37, 1029), more / fewer text rows than classes, T up to 10 (Ns = 140)."""
    from team_b200 import head
    dev = torch.device("cuda")
    C = 2 * T
    params = synth.make_params(T, seed=300 + T)
    protos = synth.make_prototypes(C, seed=11)
    image = synth.make_batch(B, C, step=40 + T)["image"]
    text = synth.make_text_class_features(20)[:Tn].contiguous()
    p64 = {k: v.double() for k, v in params.items()}
    with torch.no_grad():
        ref = O.forward_proof(p64, image.double(), text.double(), protos.double())
    pack = _pack(params, dev)
    img, txt, pro = head.forward_proof(pack, image.to(dev), text.to(dev), protos.to(dev), mode=head.MODE_F32)
    for key, o, r in (("image", img, ref[0]), ("text", txt, ref[1]), ("proto", pro, ref[3])):
        assert tuple(o.shape) == tuple(r.shape), key
        assert rel(o, r) < 1e-5, (key, rel(o, r))
    img, txt, pro = head.forward_proof(pack, image.to(dev), text.to(dev), protos.to(dev), mode=head.MODE_BF16)
    for key, o, r in (("image", img, ref[0]), ("text", txt, ref[1]), ("proto", pro, ref[3])):
        assert rel(o, r) < 1e-2, (key, rel(o, r))


def test_forward_transformer_equals_forward_on_encoded_rows():
    """forward_transformer(transformer=True) on rows encoded by encode_image / encode_text(normalize=True) is the
    same function as forward (utils/inc_net.py:465-492 vs :436-463); run-to-run results are bit-identical."""
    from team_b200 import head
    dev = torch.device("cuda")
    T, B = 2, 33
    C = 2 * T
    params = synth.make_params(T, seed=8)
    protos = synth.make_prototypes(C, seed=2).to(dev)
    image = synth.make_batch(B, C, step=3)["image"].to(dev)
    text = synth.make_text_class_features(20)[:C].contiguous().to(dev)
    pack = _pack(params, dev)
    a = head.forward_proof(pack, image, text, protos, mode=head.MODE_F32)
    a2 = head.forward_proof(pack, image, text, protos, mode=head.MODE_F32)
    for x, y in zip(a, a2):
        assert torch.equal(x, y)
    ei = head.encode(pack, "image", image, normalize=True, mode=head.MODE_F32)
    et = head.encode(pack, "text", text, normalize=True, mode=head.MODE_F32)
    b = head.forward_proof(pack, ei, et, protos, inputs_encoded=True, mode=head.MODE_F32)
    for x, y in zip(a, b):
        assert rel(x, y) < 1e-6


def test_proof_net_forward_surface(golden):
    """The drop-in class: Proof_Net.forward(image, text) returns (image, text, exp(logit_scale), proto)."""
    from test_gpu_inc_net import build_net
    case, g = CASES["proof_T2_B5"], golden("proof_T2_B5")
    ci = case_inputs(case)
    net = build_net(case["T"], ci["params"], ci["protos"])
    b = ci["batch"]
    img, txt, ls, pro = net.forward(b["image"].cuda(), b["text_cls"].cuda())
    assert rel(img, g["image"]) < 1e-5 and rel(txt, g["text"]) < 1e-5 and rel(pro, g["proto"]) < 1e-5
    assert rel(ls, g["logit_scale_exp"]) < 1e-6


def test_tri_modal_class_text_vs_golden(golden):
    """forward_tri_modal with class texts (text rows != batch): SURVEY 8a row a7, second input form."""
    from team_b200 import head
    dev = torch.device("cuda")
    case, g = CASES["head_T2_B7_classtext"], golden("head_T2_B7_classtext")
    ci = case_inputs(case)
    b = ci["batch"]
    outs = head.forward_tri_modal_class_text(_pack(ci["params"], dev), b["image"].to(dev), b["text_cls"].to(dev),
                                             b["state"].to(dev), ci["protos"].to(dev), mode=head.MODE_F32)
    for key, o in zip(("image", "text", "state", "proto"), outs):
        assert tuple(o.shape) == g[key].shape, (key, o.shape, g[key].shape)
        assert rel(o, g[key]) < 1e-5, (key, rel(o, g[key]))


@pytest.mark.parametrize("T,B,Tn,five", [(10, 100, 20, True), (3, 33, 4, False), (2, 9, 1, False)])
def test_tri_modal_class_text_vs_oracle(T, B, Tn, five):
    from team_b200 import head
    dev = torch.device("cuda")
    C = 2 * T
    params = synth.make_params(T, seed=400 + T)
    protos = synth.make_prototypes(C, seed=13)
    batch = synth.make_batch(B, C, step=50 + T, five_state=five)
    text = synth.make_text_class_features(20)[:Tn].contiguous()
    p64 = {k: v.double() for k, v in params.items()}
    with torch.no_grad():
        ref = O.forward_tri_modal(p64, batch["image"].double(), text.double(), batch["state"], protos.double())
    pack = _pack(params, dev)
    for mode, tol in ((head.MODE_F32, 1e-5), (head.MODE_BF16, 1e-2)):
        outs = head.forward_tri_modal_class_text(pack, batch["image"].to(dev), text.to(dev), batch["state"].to(dev),
                                                 protos.to(dev), mode=mode)
        for key, o, r in zip(("image", "text", "state", "proto"), outs, ref[:4]):
            assert tuple(o.shape) == tuple(r.shape), (key, o.shape, r.shape)
            assert rel(o, r) < tol, (key, mode, rel(o, r))
