"""GPU parity of the fusion head (forward_tri_modal fwd + bwd, classification logits) through
the C ABI against the CPU oracle and the reference-generated golden vectors.
fp32 mode bar: outputs and gradients <= 1e-5 relative (norm-wise), argmax exact."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs, grad_subsample

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_head(params, batch, protos, cots, mode, with_cls=True):
    from team_b200 import head
    dev = torch.device("cuda")
    p = {k: v.to(dev).requires_grad_(v.is_floating_point() and v.dim() > 0) for k, v in params.items()}
    pack = head.HeadParamPack.from_state_dict(p)
    outs = head.forward_tri_modal(pack, batch["image"].to(dev), batch["text"].to(dev), batch["state"].to(dev),
                                  protos.to(dev), text_cls=batch["text_cls"].to(dev) if with_cls else None, mode=mode)
    names = O.trainable_names(params)
    grads = torch.autograd.grad(outs[:4], [p[n] for n in names], grad_outputs=[c.to(dev) for c in cots],
                                allow_unused=True)
    return outs, dict(zip(names, grads))


@pytest.mark.parametrize("name", ["head_T1_B6", "head_T3_B5_5state", "head_T10_B4"])
def test_head_f32_vs_golden(name, golden):
    from team_b200 import head
    case, g = CASES[name], golden(name)
    ci = case_inputs(case)
    outs, grads = run_head(ci["params"], ci["batch"], ci["protos"], ci["cots"], head.MODE_F32)
    for key, o in zip(("image", "text", "state", "proto"), outs[:4]):
        assert tuple(o.shape) == g[key].shape, key
        assert rel(o, g[key]) < 1e-5, (key, rel(o, g[key]))
    assert rel(outs[4], g["cls_logits"]) < 1e-5
    assert np.array_equal(outs[5].cpu().numpy(), g["cls_logits"].argmax(1))
    for n, gr in grads.items():
        e = rel(grad_subsample(gr), g["grad:" + n])
        assert e < 2e-5, (n, e)


@pytest.mark.parametrize("T,B,five", [(1, 64, False), (10, 96, True), (4, 300, False)])
def test_head_f32_vs_oracle(T, B, five):
    from team_b200 import head
    C = 2 * T
    params = synth.make_params(T, seed=100 + T)
    protos = synth.make_prototypes(C, seed=7)
    batch = synth.make_batch(B, C, step=T, five_state=five)
    cots = synth.make_cotangents(B, step=T)
    outs, grads = run_head(params, batch, protos, cots, head.MODE_F32)
    p64 = {k: v.double().requires_grad_(v.dim() > 0) for k, v in params.items()}
    ref = O.forward_tri_modal(p64, batch["image"].double(), batch["text"].double(), batch["state"], protos.double())
    for key, o, r in zip(("image", "text", "state", "proto"), outs[:4], ref[:4]):
        assert rel(o, r) < 1e-5, (key, rel(o, r))
    names = O.trainable_names(params)
    gref = torch.autograd.grad(ref[:4], [p64[n] for n in names], grad_outputs=[c.double() for c in cots])
    for n, gr in zip(names, gref):
        assert rel(grads[n], gr) < 2e-5, (n, rel(grads[n], gr))
    logits = O.forward_for_classification({k: v.detach() for k, v in p64.items()}, batch["image"].double(),
                                          batch["text_cls"].double())
    assert rel(outs[4], logits) < 1e-5
    top2 = logits.topk(2, dim=1).values if C > 1 else None
    safe = (top2[:, 0] - top2[:, 1]) > 1e-5
    assert torch.equal(outs[5].cpu()[safe], logits.argmax(1)[safe])
    # deterministic: a second run is bit-identical
    outs2, grads2 = run_head(params, batch, protos, cots, head.MODE_F32)
    for a, b in zip(outs[:4], outs2[:4]):
        assert torch.equal(a, b)
    for n in names:
        assert torch.equal(grads[n], grads2[n]), n


def test_encode_paths():
    from team_b200 import head
    T, C = 3, 6
    params = synth.make_params(T, seed=5)
    dev = torch.device("cuda")
    pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
    batch = synth.make_batch(33, C, step=9)
    protos = synth.make_prototypes(C, seed=3)
    for norm in (False, True):
        assert rel(head.encode(pack, "image", batch["image"].to(dev), normalize=norm), O.encode_image(batch["image"], params, norm)) < 1e-5
        assert rel(head.encode(pack, "text", batch["text"].to(dev), normalize=norm), O.encode_text(batch["text"], params, norm)) < 1e-5
        assert rel(head.encode(pack, "state", batch["state"].to(dev), normalize=norm), O.encode_state(batch["state"], params, norm)) < 1e-5
        assert rel(head.encode(pack, "prototypes", None, protos.to(dev), normalize=norm), O.encode_prototypes(protos, params, norm)) < 1e-5


def _quantised_operand_params(params):
    """bf16-mode operand contract: the projections are summed in fp32 and rounded to bf16
    once; biases stay fp32.  Returned as a single-task parameter set for the oracle."""
    T = O.num_tasks(params)
    q = {}
    for kind in ("img", "text", "state"):
        W = sum(params[f"projs_{kind}.{t}.MLP.0.weight"] for t in range(T))
        b = sum(params[f"projs_{kind}.{t}.MLP.0.bias"] for t in range(T))
        q[f"projs_{kind}.0.MLP.0.weight"] = W.to(torch.bfloat16).float()
        q[f"projs_{kind}.0.MLP.0.bias"] = b
    return q


@pytest.mark.parametrize("T,B", [(1, 64), (10, 512), (10, 1024)])
def test_head_bf16_mode(T, B):
    """16-bit tensor-core mode (TEAM_MODE_BF16): every GEMM operand (inputs, weights, forward activations,
    gradients) rounded to bf16 once, fp32 accumulation in tensor memory, all row-wise math and every output fp32
    (DESIGN.md 4).  (10, 1024) is the benchmarked shape.  Bars (norm-wise relative error):
      B. against the same algorithm in fp64 with a rounding at every point where the kernels round
         (oracle/quantised_model.py, pinned on CPU against the reference-generated oracle): all four feature
         outputs <= 1e-3 (measured 3e-4), every gradient <= 5e-3 (measured <= 4e-3: the coefficient GEMMs of the
         table-row gradients round their operands too, the model evaluates them exactly);
      A. against the reference evaluated in fp64 on the identically quantised OPERANDS (only inputs / weights
         rounded - SURVEY 7.3 (i), north_star "logits within 1e-3 relative in bf16"): classification logits <= 1e-3,
         argmax exact where the fp64 margin exceeds the bound; the fused features carry one more bf16 rounding per
         GEMM stage and are held to 1e-2 here (measured 3-6e-3), gradients to 1.5e-2."""
    from team_b200 import head
    from oracle import quantised_model as Q
    C = 2 * T
    params = synth.make_params(T, seed=100 + T)
    protos = synth.make_prototypes(C, seed=7)
    batch = synth.make_batch(B, C, step=T)
    cots = synth.make_cotangents(B, step=T)
    outs, grads = run_head(params, batch, protos, cots, head.MODE_BF16)
    p64 = {k: v.double() for k, v in params.items()}
    args = (p64, batch["image"].double(), batch["text"].double(), batch["state"], protos.double(), [c.double() for c in cots])
    # ---- A: operand-quantised reference
    oa, ga, la = Q.head_fwd_bwd(*args, text_cls=batch["text_cls"].double(), operands_only=True)
    for key, o, r in zip(("image", "text", "state", "proto"), outs[:4], oa):
        assert rel(o, r) < 1e-2, ("A", key, rel(o, r))
    assert rel(outs[4], la) < 1e-3, rel(outs[4], la)
    top2 = la.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2e-3
    assert torch.equal(outs[5].cpu()[safe], la.argmax(1)[safe])
    # ---- B: rounding at the kernels' own rounding points
    ob, gb = Q.head_fwd_bwd(*args)
    for key, o, r in zip(("image", "text", "state", "proto"), outs[:4], ob):
        assert rel(o, r) < 1e-3, ("B", key, rel(o, r))
    worst = {n: rel(grads[n], gb[n]) for n in gb}
    assert max(worst.values()) < 5e-3, worst
    worst_a = {n: rel(grads[n], ga[n]) for n in ga}
    assert max(worst_a.values()) < 1.5e-2, worst_a


@pytest.mark.parametrize("mode_name,T,B", [("f32", 3, 40), ("f32", 10, 300), ("bf16", 10, 512)])
def test_gram_table_rows_match_the_oracle(mode_name, T, B, monkeypatch):
    """The Gram formulation of the table-query rows (csrc/head_table_gram.cuh, the default from 16 384 samples per
    GPU) forced on at small batches: the same parity bars as the default kernels - fp32 <= 1e-5 / 2e-5 against the
    fp64 oracle, bf16 against the model that rounds where the kernels round - and it must agree with the
    second-generation kernels on every output and gradient."""
    from team_b200 import head
    from oracle import quantised_model as Q
    C = 2 * T
    params = synth.make_params(T, seed=300 + T)
    protos = synth.make_prototypes(C, seed=11)
    batch = synth.make_batch(B, C, step=T, five_state=(T == 3))
    cots = synth.make_cotangents(B, step=T)
    mode = head.MODE_F32 if mode_name == "f32" else head.MODE_BF16
    monkeypatch.setenv("TEAM_TABLE_GRAM_MIN_B", "0")
    outs_g, grads_g = run_head(params, batch, protos, cots, mode)
    monkeypatch.setenv("TEAM_TABLE_GRAM_MIN_B", "1000000000")
    outs_2, grads_2 = run_head(params, batch, protos, cots, mode)
    tol = 2e-5 if mode_name == "f32" else 2e-3
    for a, b in zip(outs_g[:4], outs_2[:4]):
        assert rel(a, b) < tol, rel(a, b)
    assert max(rel(grads_g[n], grads_2[n]) for n in grads_2) < (5e-5 if mode_name == "f32" else 5e-3)
    p64 = {k: v.double().requires_grad_(v.dim() > 0) for k, v in params.items()}
    if mode_name == "f32":
        ref = O.forward_tri_modal(p64, batch["image"].double(), batch["text"].double(), batch["state"], protos.double())
        names = O.trainable_names(params)
        gref = torch.autograd.grad(ref[:4], [p64[n] for n in names], grad_outputs=[c.double() for c in cots])
        for o, r in zip(outs_g[:4], ref[:4]):
            assert rel(o, r) < 1e-5, rel(o, r)
        for n, gr in zip(names, gref):
            assert rel(grads_g[n], gr) < 2e-5, (n, rel(grads_g[n], gr))
    else:
        pd = {k: v.double() for k, v in params.items()}
        ob, gb = Q.head_fwd_bwd(pd, batch["image"].double(), batch["text"].double(), batch["state"], protos.double(), [c.double() for c in cots])
        for o, r in zip(outs_g[:4], ob):
            assert rel(o, r) < 1e-3, rel(o, r)
        assert max(rel(grads_g[n], gb[n]) for n in gb) < 5e-3


def test_host_batch_pipeline_matches_direct_step():
    """HostBatchPipeline (pinned host batches, copy stream, graph replay, D2H predictions) gives the same
    predictions and the same gradient bucket as a direct HeadStepRunner step on the same data."""
    from team_b200 import head
    dev = torch.device("cuda")
    T, B = 3, 40
    C = 2 * T
    params = synth.make_params(T, seed=5)
    pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
    protos = synth.make_prototypes(C, seed=3).to(dev)
    text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
    batches = [synth.make_batch(B, C, step=s) for s in range(5)]
    cots = [[c.reshape(B, 512) for c in synth.make_cotangents(B, step=s)] for s in range(5)]
    ref = head.HeadStepRunner(pack, protos, B, C, head.MODE_F32)
    want_pred, want_grad = [], []
    for b, c in zip(batches, cots):
        ref.step(b["image"].to(dev), b["text"].to(dev), b["state"].to(dev), text_cls, [x.to(dev) for x in c])
        torch.cuda.synchronize()
        want_pred.append(ref.argmax.cpu().clone()); want_grad.append(ref.flat_grads.cpu().clone())
    seen_grads = []
    pipe = head.HostBatchPipeline(pack, protos, B, text_cls, mode=head.MODE_F32, depth=2,
                                  after_step=lambda r: seen_grads.append(r.flat_grads.clone()))
    got = []
    for b, c in zip(batches, cots):
        out = pipe.submit(b["image"].pin_memory(), b["text"].pin_memory(), b["state"].pin_memory(),
                          [x.pin_memory() for x in c])
        if out is not None:
            got.append(out)
    got += pipe.drain()
    assert len(got) == 5
    assert pipe.h2d_bytes_per_step == B * 512 * 4 * 6 + B * 8 and pipe.d2h_bytes_per_step == B * 8
    for g, w in zip(got, want_pred):
        assert torch.equal(g, w)
    torch.cuda.synchronize()
    for g, w in zip(seen_grads, want_grad):
        assert torch.equal(g.cpu(), w)          # same kernels, same order: bit-identical


def test_grad_ready_events_inside_a_captured_step():
    """team_head_grads.ev_*: the events recorded in the middle of a CAPTURED backward (external record nodes)
    release a side stream that reads the finished gradient buckets while the rest of the graph still runs."""
    from team_b200 import head
    dev = torch.device("cuda")
    T, B = 2, 48
    C = 2 * T
    params = synth.make_params(T, seed=9)
    pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
    protos = synth.make_prototypes(C, seed=4).to(dev)
    text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
    b = synth.make_batch(B, C, step=1)
    args = (b["image"].to(dev), b["text"].to(dev), b["state"].to(dev), text_cls,
            [c.reshape(B, 512).to(dev) for c in synth.make_cotangents(B, step=1)])
    plain = head.HeadStepRunner(pack, protos, B, C, head.MODE_F32)
    plain.step(*args)
    torch.cuda.synchronize()
    want = plain.flat_grads.clone()
    r = head.HeadStepRunner(pack, protos, B, C, head.MODE_F32, grad_events=True)
    assert sum(x.numel() for x in r.buckets) == r.flat_grads.numel()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        r.step(*args)                       # eager: plain event records
        st.synchronize()
        assert torch.equal(r.flat_grads, want)
        r.flat_grads.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            r.step(*args)
        r.flat_grads.zero_()
        st.synchronize()
        g.replay()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        side.wait_event(r.ready_events[0])
        got0 = r.buckets[0].clone()
        side.wait_event(r.ready_events[1])
        got1 = r.buckets[1].clone()
    torch.cuda.synchronize()
    d2 = 512 * 512
    assert torch.equal(got0, want[:d2]) and torch.equal(got1, want[d2:4 * d2])
    assert torch.equal(r.flat_grads, want)


@pytest.mark.parametrize("mode_name,B,chunks", [("bf16", 4096, 8), ("f32", 2048, 4)])
def test_large_batch_equals_its_chunks(mode_name, B, chunks):
    """Full-size batches take other code paths than the small parity cases (persistent 128x256 GEMM + split-K
    fix-up kernel, several rounds per CTA in the table-row kernels).  Size-independent properties pin them:
      * every output row depends on its own sample only  -> one call on B samples == `chunks` calls on B/chunks,
      * the parameter gradients are sums over samples     -> grad(B) == sum of the chunk gradients,
      * the first 48 samples agree with the fp64 oracle run on those 48 samples alone."""
    from team_b200 import head
    mode = head.MODE_BF16 if mode_name == "bf16" else head.MODE_F32
    dev = torch.device("cuda")
    T = 10
    C = 2 * T
    params = synth.make_params(T, seed=77)
    pack = head.HeadParamPack.from_state_dict({k: v.to(dev) for k, v in params.items()})
    protos = synth.make_prototypes(C, seed=5)
    batch = synth.make_batch(B, C, step=9, five_state=True)
    cots = [c.reshape(B, 512) for c in synth.make_cotangents(B, step=9)]
    text_cls = batch["text_cls"].to(dev)
    img, txt, sid = batch["image"].to(dev), batch["text"].to(dev), batch["state"].to(dev)
    cd = [c.to(dev) for c in cots]
    full = head.HeadStepRunner(pack, protos.to(dev), B, C, mode)
    full.step(img, txt, sid, text_cls, cd)
    torch.cuda.synchronize()
    outs_full, grad_full, am_full = full.outs.clone(), full.flat_grads.clone(), full.argmax.clone()
    n = B // chunks
    part = head.HeadStepRunner(pack, protos.to(dev), n, C, mode)
    grad_sum = torch.zeros_like(grad_full, dtype=torch.float64)
    tol_o, tol_g = (2e-3, 4e-3) if mode_name == "bf16" else (1e-5, 3e-5)
    for c in range(chunks):
        sl = slice(c * n, (c + 1) * n)
        part.step(img[sl].contiguous(), txt[sl].contiguous(), sid[sl].contiguous(), text_cls, [x[sl].contiguous() for x in cd])
        torch.cuda.synchronize()
        for k in range(4):
            assert rel(outs_full[k, sl], part.outs[k]) < tol_o, (c, k, rel(outs_full[k, sl], part.outs[k]))
        assert (am_full[sl] == part.argmax).float().mean() > 0.999
        grad_sum += part.flat_grads.double()
    for name, view in full.grad_views.items():
        lo = view.data_ptr() - full.flat_grads.data_ptr()
        seg = slice(lo // 4, lo // 4 + view.numel())
        e = rel(grad_full[seg], grad_sum[seg])
        assert e < tol_g, (name, e)
    # oracle on the first samples alone
    m = 48
    p64 = {k: v.double() for k, v in params.items()}
    with torch.no_grad():
        ref = O.forward_tri_modal(p64, batch["image"][:m].double(), batch["text"][:m].double(), batch["state"][:m], protos.double())
    tol = 1e-2 if mode_name == "bf16" else 1e-5
    for k, r in enumerate(ref[:4]):
        assert rel(outs_full[k, :m], r.reshape(m, 512)) < tol, (k, rel(outs_full[k, :m], r.reshape(m, 512)))


@pytest.mark.parametrize("mode_name", ["f32", "bf16"])
def test_null_proto_cotangent_equals_zero_cotangent(mode_name):
    """g_proto = NULL (the learner's losses never read the prototype output) skips the C prototype rows of every
    sample in the backward; the gradients must equal those of an explicit all-zero cotangent, and through autograd an
    unused prototype output takes the same path."""
    from team_b200 import head
    mode = head.MODE_BF16 if mode_name == "bf16" else head.MODE_F32
    dev = torch.device("cuda")
    T, B = 4, 70
    C = 2 * T
    params = synth.make_params(T, seed=31)
    pdev = {k: v.to(dev) for k, v in params.items()}
    pack = head.HeadParamPack.from_state_dict(pdev)
    protos = synth.make_prototypes(C, seed=8).to(dev)
    text_cls = synth.make_text_class_features(20)[:C].contiguous().to(dev)
    b = synth.make_batch(B, C, step=6, five_state=True)
    img, txt, sid = b["image"].to(dev), b["text"].to(dev), b["state"].to(dev)
    cots = [c.reshape(B, 512).to(dev) for c in synth.make_cotangents(B, step=6)]
    r = head.HeadStepRunner(pack, protos, B, C, mode)
    r.forward(img, txt, sid, text_cls)
    r.backward(img, txt, sid, [cots[0], cots[1], cots[2], torch.zeros_like(cots[3])])
    torch.cuda.synchronize()
    want = r.flat_grads.clone()
    r.forward(img, txt, sid, text_cls)
    r.backward(img, txt, sid, [cots[0], cots[1], cots[2], None])
    torch.cuda.synchronize()
    assert rel(r.flat_grads, want) < 1e-6, rel(r.flat_grads, want)
    # autograd: a loss that ignores the prototype output
    names = O.trainable_names(params)
    p = {k: v.clone().requires_grad_(k in names) for k, v in pdev.items()}
    outs = head.forward_tri_modal(head.HeadParamPack.from_state_dict(p), img, txt, sid, protos, mode=mode)
    loss = (outs[0] * cots[0]).sum() + (outs[1].reshape(B, 512) * cots[1]).sum() + (outs[2] * cots[2]).sum()
    grads = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
    got = {n: g for n, g in zip(names, grads) if g is not None}
    for name, view in r.grad_views.items():
        key = {"w_q": "sel_attn.w_qs.weight", "w_fc": "sel_attn.fc.weight", "state_emb": "state_embedder.state_embeddings.weight",
               "w_img": f"projs_img.{T - 1}.MLP.0.weight"}.get(name)
        if key is not None:
            lo = (view.data_ptr() - r.flat_grads.data_ptr()) // 4
            assert rel(got[key].reshape(-1), want[lo:lo + view.numel()]) < 1e-6, name


def test_frozen_projection_sums_are_bit_identical():
    """team_head_weights.num_frozen / team_head_frozen_sums: the old tasks' projections enter the step prologue as one
    precomputed term (they are frozen, utils/inc_net.py:392-393).  Outputs and gradients must equal the all-tasks sum bit for
    bit (same left-to-right order), and an in-place change of an old projection must be picked up."""
    from team_b200 import head
    dev = torch.device("cuda")
    T, B = 4, 37
    C = 2 * T
    params = {k: v.to(dev) for k, v in synth.make_params(T, seed=91).items()}
    pack = head.HeadParamPack.from_state_dict(params)
    protos = synth.make_prototypes(C).to(dev)
    b = {k: v.to(dev) for k, v in synth.make_batch(B, C, step=5).items()}
    cots = [c.to(dev).reshape(B, 512) for c in synth.make_cotangents(B, step=5)]
    tc = synth.make_text_class_features(20)[:C].contiguous().to(dev)
    for mode in (head.MODE_F32, head.MODE_BF16):
        cached = head.HeadStepRunner(pack, protos, B, C, mode)
        assert cached.hw.num_frozen == T - 1
        plain = head.HeadStepRunner(pack, protos, B, C, mode)
        plain.hw.num_frozen = 0
        for r in (cached, plain):
            r.step(b["image"], b["text"], b["state"], tc, cots)
        torch.cuda.synchronize()
        assert torch.equal(cached.outs, plain.outs) and torch.equal(cached.flat_grads, plain.flat_grads)
        assert torch.equal(cached.logits, plain.logits)
        params["projs_img.0.MLP.0.weight"].mul_(1.5)          # e.g. load_state_dict into a frozen projection
        for r in (cached, plain):
            r.step(b["image"], b["text"], b["state"], tc, cots)
        torch.cuda.synchronize()
        assert torch.equal(cached.outs, plain.outs) and torch.equal(cached.flat_grads, plain.flat_grads)
        params["projs_img.0.MLP.0.weight"].div_(1.5)
    params["projs_text.1.MLP.0.bias"].requires_grad_(True)     # an old task that is trainable: no caching
    assert head.HeadStepRunner(pack, protos, B, C, head.MODE_F32).hw.num_frozen == 0


def test_extra_cotangent_on_own_rows():
    """team_head_grads.g_own_rows / team_head_own_rows_offset: the forward leaves normalize(encode_image(x)) |
    normalize(encode_text(t)) in its workspace, and a cotangent on those rows joins the backward - equal to running the two
    encodes and their backward separately (the ClipLoss branch of the training step, models/proof.py:428-431)."""
    from team_b200 import head
    dev = torch.device("cuda")
    T, B = 3, 29
    C = 2 * T
    params = {k: v.to(dev) for k, v in synth.make_params(T, seed=92).items()}
    pack = head.HeadParamPack.from_state_dict(params)
    protos = synth.make_prototypes(C).to(dev)
    b = {k: v.to(dev) for k, v in synth.make_batch(B, C, step=6).items()}
    cots = [c.to(dev).reshape(B, 512) for c in synth.make_cotangents(B, step=6)]
    tc = synth.make_text_class_features(20)[:C].contiguous().to(dev)
    g = torch.randn(2, B, 512, generator=torch.Generator().manual_seed(3)).to(dev)
    for mode, tol in ((head.MODE_F32, 2e-5), (head.MODE_BF16, 5e-3)):
        r = head.HeadStepRunner(pack, protos, B, C, mode)
        r.forward(b["image"], b["text"], b["state"], tc)
        xo = r.own_rows().clone()
        ei = head.encode(pack, "image", b["image"], normalize=True, mode=mode)
        et = head.encode(pack, "text", b["text"], normalize=True, mode=mode)
        assert rel(xo[:B], ei) < 1e-6 and rel(xo[B:], et) < 1e-6
        r.backward(b["image"], b["text"], b["state"], cots)
        base = {k: v.clone() for k, v in r.grad_views.items()}
        r.forward(b["image"], b["text"], b["state"], tc)
        r.backward(b["image"], b["text"], b["state"], cots, g_own_rows=g)
        torch.cuda.synchronize()
        leaf = {n: params[f"projs_{n}.{T - 1}.MLP.0.{w}"].clone().requires_grad_(True) for n in ("img", "text") for w in ("weight",)}
        p2 = dict(params)
        names = [f"projs_img.{T - 1}.MLP.0.weight", f"projs_img.{T - 1}.MLP.0.bias", f"projs_text.{T - 1}.MLP.0.weight", f"projs_text.{T - 1}.MLP.0.bias"]
        for n in names:
            p2[n] = params[n].clone().requires_grad_(True)
        pk2 = head.HeadParamPack.from_state_dict(p2)
        yi = head.encode_grad(pk2, "image", b["image"], normalize=True, mode=mode)
        yt = head.encode_grad(pk2, "text", b["text"], normalize=True, mode=mode)
        extra = torch.autograd.grad([yi, yt], [p2[n] for n in names], grad_outputs=[g[0], g[1]])
        for key, e in zip(("w_img", "b_img", "w_text", "b_text"), extra):
            if mode == head.MODE_F32:          # the increment itself
                got = r.grad_views[key] - base[key]
                assert rel(got.reshape(e.shape), e) < tol, (mode, key, rel(got.reshape(e.shape), e))
            else:                              # bf16 operands round the SUM of the cotangents: compare the totals
                want = base[key].reshape(e.shape) + e
                assert rel(r.grad_views[key].reshape(e.shape), want) < tol, (mode, key, rel(r.grad_views[key].reshape(e.shape), want))
        for key in ("w_q", "w_fc", "ln_g", "state_emb", "prompts"):
            assert torch.equal(r.grad_views[key], base[key]), key          # nothing else sees the extra cotangent
