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
    """net.sel_attn(q, k, v) as a module call (the reference's own interface), eval mode."""
    from team_b200 import inc_net
    dev = torch.device("cuda")
    m = inc_net.MultiHeadAttention(1, 512, 512, 512, dropout=0.1).to(dev)
    x = torch.randn(4, 6, 512, device=dev)
    m.eval()
    y = m(x, x, x)
    p = {"sel_attn." + n: t.detach().double().cpu() for n, t in m.named_parameters()}
    assert rel(y, O.sel_attn(x.double().cpu(), p)) < 1e-5
    y.sum().backward()
    assert all(t.grad is not None for t in m.parameters())


@pytest.mark.parametrize("B,Lq,Lk,pd", [(3, 5, 7, 0.1), (2, 141, 141, 0.1), (4, 17, 33, 0.5)])
def test_mha_train_mode_dropout_vs_oracle(B, Lq, Lk, pd):
    """Train mode: both dropouts of the block (convs/projections.py:35, :84) with the library's Philox masks, against the
    reference math with the SAME masks (oracle.philox_keep_mask, an independent numpy Philox4x32-10): outputs 1e-5, every
    input and parameter gradient 2e-5 - i.e. the masks are applied where the reference applies them, scaled by 1 / (1 - p),
    and regenerated identically in the backward."""
    from team_b200 import head
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(B * 100 + Lq)
    params = synth.make_params(1, seed=70 + B)
    q, k, v, cot = (torch.randn(s, generator=gen) for s in ((B, Lq, 512), (B, Lk, 512), (B, Lk, 512), (B, Lq, 512)))
    seed, offset = 0x1234567890ABCDEF, 2 * (2 ** 40 + 77)
    p64 = {n: t.double().requires_grad_(True) for n, t in params.items() if n.startswith("sel_attn.")}
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
    ref = O.mha(q64, k64, v64, p64, drop=(pd, seed, offset))
    gref = torch.autograd.grad(ref, [q64, k64, v64] + [p64["sel_attn." + n] for n in MHA_NAMES], grad_outputs=cot.double())
    par = [params["sel_attn." + n].to(dev).requires_grad_(True) for n in MHA_NAMES]
    qd, kd, vd = (t.to(dev).requires_grad_(True) for t in (q, k, v))
    out = head.mha(qd, kd, vd, *par, mode=head.MODE_F32, dropout_p=pd, seed=seed, offset=offset)
    grads = torch.autograd.grad(out, [qd, kd, vd] + par, grad_outputs=cot.to(dev))
    assert rel(out, ref) < 1e-5, rel(out, ref)
    for i, (a, b) in enumerate(zip(grads, gref)):
        assert rel(a, b) < 2e-5, (i, rel(a, b))
    # eval output differs (dropout really happened), another offset gives another mask
    ev = head.mha(qd, kd, vd, *par, mode=head.MODE_F32)
    other = head.mha(qd, kd, vd, *par, mode=head.MODE_F32, dropout_p=pd, seed=seed, offset=offset + 2)
    assert rel(out, ev) > 1e-3 and rel(out, other) > 1e-3
    again = head.mha(qd, kd, vd, *par, mode=head.MODE_F32, dropout_p=pd, seed=seed, offset=offset)
    assert torch.equal(again, out)


def test_dropout_mask_stream_statistics():
    """The mask stream: equals the numpy Philox bit for bit, keep rate p within 4 sigma, no correlation between neighbours,
    between consecutive offsets or between seeds."""
    from team_b200 import head
    n = 1 << 20
    for pd in (0.1, 0.5):
        m = head.dropout_keep_mask(n, pd, seed=42, offset=6).cpu().double()
        assert torch.equal(m, O.philox_keep_mask(n, pd, 42, 6))
        sigma = (pd * (1 - pd) / n) ** 0.5
        assert abs(float(m.mean()) - (1 - pd)) < 4 * sigma
        for other in (head.dropout_keep_mask(n, pd, seed=42, offset=7).cpu().double(),
                      head.dropout_keep_mask(n, pd, seed=43, offset=6).cpu().double(), torch.roll(m, 1)):
            corr = float(((m - m.mean()) * (other - other.mean())).mean() / (m.var() * other.var()).sqrt())
            assert abs(corr) < 5 / n ** 0.5, corr


def test_proof_net_train_mode_uses_dropout():
    """Proof_Net in train mode with the reference's default p = 0.1: forward_tri_modal / forward run the token-tensor route with
    dropout (no exception, outputs differ from eval mode and between calls, gradients flow); torch.manual_seed reproduces it;
    p = 0 in train mode is the factorised path and equals eval mode."""
    dev = torch.device("cuda")
    ci = case_inputs(CASES["head_T2_B7_classtext"])
    net = _net(ci["params"], ci["protos"], dev)
    b = ci["batch"]
    a = (b["image"].to(dev), b["text"].to(dev), b["state"].to(dev))
    with torch.no_grad():
        ev = net.forward_tri_modal(*a)
    net.train()
    torch.manual_seed(5)
    t1 = net.forward_tri_modal(*a)
    t2 = net.forward_tri_modal(*a)
    torch.manual_seed(5)
    t3 = net.forward_tri_modal(*a)
    assert t1[1].shape == ev[1].shape == (7, 1, 512)
    assert rel(t1[0], ev[0]) > 1e-3 and rel(t1[0], t2[0]) > 1e-3 and torch.equal(t1[0], t3[0])
    loss = sum(o.square().sum() for o in t1[:4])
    loss.backward()
    assert net.sel_attn.w_qs.weight.grad is not None and net.projs_img[-1].MLP[0].weight.grad is not None
    pf = net.forward(a[0], b["text_cls"].to(dev))
    assert pf[0].requires_grad
    net.sel_attn.dropout.p = 0.0
    t0 = net.forward_tri_modal(*a)
    for x, y in zip(t0[:4], ev[:4]):
        assert rel(x, y) < 1e-6
