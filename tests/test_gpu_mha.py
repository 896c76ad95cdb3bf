"""GPU parity of the standalone MultiHeadAttention op (team_mha_fwd / team_mha_bwd; convs/projections.py:64-87, SURVEY 8a row
a6) and of everything composed from it: the differentiable PROOF fusion (utils/inc_net.py:436-492, row a8), the class-text
form of forward_tri_modal with autograd (:544-547, :573-576) and InsectLifecycleModel.get_state_embeddings (row a5).
fp32 mode bar: <= 1e-5 outputs / 2e-5 gradients (norm-wise) against the golden vectors of the real reference and the fp64
oracle; bf16 mode (dense projections on tcgen05 with bf16 operands, attention core fp32): <= 1e-2."""
import pytest
import torch

from oracle import synth
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs, grad_subsample

pytestmark = pytest.mark.gpu

MHA_NAMES = ("w_qs.weight", "w_ks.weight", "w_vs.weight", "fc.weight", "fc.bias", "layer_norm.weight", "layer_norm.bias")


def rel(a, b):
    a = a.detach().double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _mha_gpu(ci_params, q, k, v, cot, mode, dev):
    from team_b200 import head
    par = [ci_params["sel_attn." + n].to(dev).requires_grad_(True) for n in MHA_NAMES]
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    out = head.mha(qd, kd, vd, *par, mode=mode)
    grads = torch.autograd.grad(out, [qd, kd, vd] + par, grad_outputs=cot.to(dev))
    return out, grads


def test_mha_vs_golden(golden):
    from team_b200 import head
    dev = torch.device("cuda")
    case, g = CASES["mha_cross"], golden("mha_cross")
    ci = case_inputs(case)
    out, grads = _mha_gpu(ci["params"], ci["q"], ci["k"], ci["v"], ci["cot"], head.MODE_F32, dev)
    assert rel(out, g["out"]) < 1e-5
    for n, gr in zip(("q", "k", "v", "w_q", "w_k", "w_v", "w_fc", "b_fc", "ln_g", "ln_b"), grads):
        assert rel(grad_subsample(gr), g["grad:" + n]) < 2e-5, (n, rel(grad_subsample(gr), g["grad:" + n]))


@pytest.mark.parametrize("B,Lq,Lk", [(1, 1, 1), (2, 141, 141), (5, 33, 70), (70, 9, 15)])
def test_mha_vs_oracle(B, Lq, Lk):
    """Ragged lengths (141 = the PROOF token count at T = 10), single token, more samples than one grid row."""
    from team_b200 import head
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(B * 1000 + Lq)
    params = synth.make_params(1, seed=60 + B)
    q, k, v, cot = (torch.randn(s, generator=gen) for s in ((B, Lq, 512), (B, Lk, 512), (B, Lk, 512), (B, Lq, 512)))
    p64 = {n: t.double().requires_grad_(True) for n, t in params.items() if n.startswith("sel_attn.")}
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = O.mha(q64, k64, v64, p64)
    gref = torch.autograd.grad(ref, [q64, k64, v64] + [p64["sel_attn." + n] for n in MHA_NAMES], grad_outputs=cot.double())
    for mode, to, tg in ((head.MODE_F32, 1e-5, 2e-5), (head.MODE_BF16, 1e-2, 2e-2)):
        out, grads = _mha_gpu(params, q, k, v, cot, mode, dev)
        assert rel(out, ref) < to, (mode, rel(out, ref))
        for i, (a, b) in enumerate(zip(grads, gref)):
            assert rel(a, b) < tg, (mode, i, rel(a, b))
    o2, g2 = _mha_gpu(params, q, k, v, cot, head.MODE_F32, dev)          # run-to-run reproducible
    o3, g3 = _mha_gpu(params, q, k, v, cot, head.MODE_F32, dev)
    assert torch.equal(o2, o3) and all(torch.equal(a, b) for a, b in zip(g2, g3))


def _net(params, protos, dev, mode="f32"):
    from team_b200 import inc_net
    T = O.num_tasks(params)
    args = {"device": [dev], "projection_type": "pure_mlp", "context_prompt_length_per_task": params["context_prompts.0"].shape[0],
            "team_mode": mode}

    class _Clip(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.logit_scale = torch.nn.Parameter(params["convnet.logit_scale"].clone())

        def encode_image(self, x, normalize=False):
            return x

        def encode_text(self, x, normalize=False):
            return x

    net = inc_net.Proof_Net(args, False, convnet=_Clip(), tokenizer=lambda t: t)
    for t in range(T):
        net.update_prototype(2 * (t + 1)); net.update_context_prompt(); net.extend_task()
    sd = net.state_dict()
    with torch.no_grad():
        for k_, v_ in params.items():
            sd[k_].copy_(v_)
    net.img_prototypes = protos.clone().to(dev)
    net.to(dev)
    net.freeze_projection_weight_new()
    net.eval()
    return net


def test_proof_forward_autograd_vs_golden(golden):
    """Proof_Net.forward under enable_grad: outputs and every trainable-parameter gradient of the REAL reference."""
    dev = torch.device("cuda")
    case, g = CASES["proof_T2_B5_grad"], golden("proof_T2_B5_grad")
    ci = case_inputs(case)
    net = _net(ci["params"], ci["protos"], dev)
    b = ci["batch"]
    img, txt, ls, pr = net.forward(b["image"].to(dev), b["text_cls"].to(dev))
    assert img.requires_grad and txt.requires_grad and pr.requires_grad
    for key, o in (("image", img), ("text", txt), ("proto", pr)):
        assert tuple(o.shape) == g[key].shape and rel(o, g[key]) < 1e-5, (key, rel(o, g[key]))
    cots = [ci["cots"][0].to(dev), ci["cots"][2][:txt.shape[0]].to(dev), ci["cots"][3][:ci["C"]].to(dev)]
    names = O.trainable_names(ci["params"])
    sd = dict(net.named_parameters())
    grads = torch.autograd.grad([img, txt, pr], [sd[n] for n in names], grad_outputs=cots, allow_unused=True)
    for n, gr in zip(names, grads):
        want = g["grad:" + n]
        if gr is None:
            assert want.size == 0 or not want.any(), n
        else:
            assert rel(grad_subsample(gr), want) < 2e-5, (n, rel(grad_subsample(gr), want))
    # the no-grad call takes the fused forward-only kernel and returns the same values
    with torch.no_grad():
        f = net.forward(b["image"].to(dev), b["text_cls"].to(dev))
    assert not f[0].requires_grad
    for a, b_ in zip((img, txt, pr), (f[0], f[1], f[3])):
        assert rel(a, b_) < 1e-5
    # forward_transformer(transformer=True) on encoded rows is the same differentiable function
    ei = net.encode_image(b["image"].to(dev), normalize=True)
    et = net.encode_text(b["text_cls"].to(dev), normalize=True)
    t = net.forward_transformer(ei, et, transformer=True)
    for a, b_ in zip((img, txt, pr), (t[0], t[1], t[3])):
        assert rel(b_, a) < 1e-6
    g2 = torch.autograd.grad([t[0], t[1], t[3]], [sd[n] for n in names], grad_outputs=cots, allow_unused=True)
    for n, a, b_ in zip(names, grads, g2):
        assert (a is None) == (b_ is None) and (a is None or rel(b_, a) < 1e-5), n


def test_class_text_autograd_vs_golden(golden):
    """forward_tri_modal with class texts (text rows != batch) under enable_grad: the real reference's outputs and gradients."""
    dev = torch.device("cuda")
    case, g = CASES["head_T2_B7_classtext"], golden("head_T2_B7_classtext")
    ci = case_inputs(case)
    net = _net(ci["params"], ci["protos"], dev)
    b = ci["batch"]
    outs = net.forward_tri_modal(b["image"].to(dev), b["text_cls"].to(dev), b["state"].to(dev))
    for key, o in zip(("image", "text", "state", "proto"), outs[:4]):
        assert tuple(o.shape) == g[key].shape and rel(o, g[key]) < 1e-5, (key, rel(o, g[key]))
    cots = [c.to(dev) for c in ci["cots"]]
    cots[1] = cots[1].view(cots[1].shape[0], -1)
    names = O.trainable_names(ci["params"])
    sd = dict(net.named_parameters())
    grads = torch.autograd.grad(outs[:4], [sd[n] for n in names], grad_outputs=cots, allow_unused=True)
    for n, gr in zip(names, grads):
        assert rel(grad_subsample(gr), g["grad:" + n]) < 2e-5, (n, rel(grad_subsample(gr), g["grad:" + n]))
    with torch.no_grad():                       # fused forward-only kernel: same values
        f = net.forward_tri_modal(b["image"].to(dev), b["text_cls"].to(dev), b["state"].to(dev))
    for a, b_ in zip(outs[:4], f[:4]):
        assert rel(b_, a) < 1e-5


def test_state_embedding_lookup():
    """InsectLifecycleModel.get_state_embeddings (models/state_evolution.py:45-47): gather + deterministic scatter-sum backward."""
    from team_b200 import inc_net
    dev = torch.device("cuda")
    m = inc_net.InsectLifecycleModel(512, 256, 10).to(dev)
    ids = torch.tensor([[1, 4, 4], [3, 1, 9]], device=dev)
    out = m.get_state_embeddings(ids)
    assert out.shape == (2, 3, 512) and torch.equal(out, m.state_embeddings.weight.detach()[ids])
    cot = torch.randn(2, 3, 512, device=dev)
    (gw,) = torch.autograd.grad(out, [m.state_embeddings.weight], grad_outputs=cot)
    want = torch.zeros(10, 512, dtype=torch.float64)
    want.index_add_(0, ids.reshape(-1).cpu(), cot.reshape(-1, 512).double().cpu())
    assert rel(gw, want) < 1e-6 and not gw[0].any()


def test_standalone_sel_attn_module():
    """net.sel_attn(q, k, v) as a module call (the reference's own interface), train mode with p = 0."""
    from team_b200 import inc_net
    dev = torch.device("cuda")
    m = inc_net.MultiHeadAttention(1, 512, 512, 512, dropout=0.1).to(dev)
    x = torch.randn(4, 6, 512, device=dev)
    m.train()
    with pytest.raises(NotImplementedError):
        m(x, x, x)
    m.dropout.p = 0.0
    y = m(x, x, x)
    p = {"sel_attn." + n: t.detach().double().cpu() for n, t in m.named_parameters()}
    assert rel(y, O.sel_attn(x.double().cpu(), p)) < 1e-5
    y.sum().backward()
    assert all(t.grad is not None for t in m.parameters())
