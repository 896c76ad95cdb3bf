"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sample sharding, the
prototype-sum all-reduce (counts exact, means equal to the single-process oracle) and the
gradient-bucket all-reduce (sum over shards == full-batch gradient of a batch-summed loss)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth
from oracle import team_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        ret[rank] = fn(rank, world_size)
    finally:
        dist.destroy_process_group()


def _run(fn, world_size=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world_size, _free_port(), fn, ret), nprocs=world_size, join=True)
    return [ret[r] for r in range(world_size)]


def test_shard_bounds_cover_every_row_once():
    from team_b200 import parallel
    for n in (0, 1, 7, 64, 1023, 4096):
        for ws in (1, 2, 3, 8):
            seen = []
            for r in range(ws):
                b, e = parallel.shard_bounds(n, r, ws)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def _proto_job(rank, ws):
    from team_b200 import parallel
    x, y, s = synth.make_prototype_build_inputs(1001, 6, seed=3)     # ragged: 1001 rows over 2 ranks
    b, e = parallel.shard_bounds(x.shape[0], rank, ws)
    K = 6 * 10
    key = y[b:e] * 10 + s[b:e]
    sums = torch.zeros(K, 512).index_add_(0, key, x[b:e])            # stands in for team_segsum on this shard
    counts = torch.bincount(key, minlength=K)
    parallel.allreduce_prototype_sums(sums, counts)
    return sums, counts


def test_prototype_sum_allreduce_matches_single_process():
    outs = _run(_proto_job)
    x, y, s = synth.make_prototype_build_inputs(1001, 6, seed=3)
    key = y * 10 + s
    ref_counts = torch.bincount(key, minlength=60)
    ref_sums = torch.zeros(60, 512, dtype=torch.float64).index_add_(0, key, x.double())
    for sums, counts in outs:
        assert torch.equal(counts, ref_counts)                       # integer counts: exact
        assert float((sums.double() - ref_sums).abs().max()) < 1e-4
    assert torch.equal(outs[0][0], outs[1][0])                       # every rank holds the same bits
    # means from the reduced sums == the oracle's class prototypes (models/proof.py:258-276)
    cls_sums = outs[0][0].view(6, 10, 512).sum(1)
    cls_counts = outs[0][1].view(6, 10).sum(1)
    means = cls_sums / cls_counts.clamp_min(1).unsqueeze(1)
    ref = O.simplecil_prototypes(x, y, torch.zeros(6, 512))
    assert float((means - ref).abs().max()) < 1e-5


def _grad_job(rank, ws):
    from team_b200 import parallel
    T, B = 2, 12
    C = 2 * T
    params = synth.make_params(T, seed=5)
    names = O.trainable_names(params)
    p = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}
    protos = synth.make_prototypes(C)
    batch = synth.make_batch(B, C, step=1)
    cots = synth.make_cotangents(B, step=1)
    b, e = parallel.shard_bounds(B, rank, ws)
    outs = O.forward_tri_modal(p, batch["image"][b:e], batch["text"][b:e], batch["state"][b:e], protos)
    g = torch.autograd.grad(outs[:4], [p[n] for n in names], grad_outputs=[c[b:e] for c in cots])
    flat = torch.cat([x.reshape(-1) for x in g])                      # the flat bucket of HeadStepRunner
    parallel.allreduce_gradients(flat)
    return flat


def test_gradient_allreduce_equals_full_batch_gradient():
    outs = _run(_grad_job)
    T, B = 2, 12
    C = 2 * T
    params = synth.make_params(T, seed=5)
    names = O.trainable_names(params)
    p = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}
    protos = synth.make_prototypes(C)
    batch = synth.make_batch(B, C, step=1)
    cots = synth.make_cotangents(B, step=1)
    ref_outs = O.forward_tri_modal(p, batch["image"], batch["text"], batch["state"], protos)
    g = torch.autograd.grad(ref_outs[:4], [p[n] for n in names], grad_outputs=list(cots))
    ref = torch.cat([x.reshape(-1) for x in g])
    for flat in outs:
        assert float((flat - ref).norm() / ref.norm()) < 1e-5
    assert torch.equal(outs[0], outs[1])


def _means_job(rank, ws):
    from team_b200 import parallel
    g = torch.Generator().manual_seed(9)
    rows = torch.randn(10, 3, 512, generator=g)
    b, e = parallel.shard_bounds(10, rank, ws)
    t, p = parallel.allreduce_batch_means(rows[b:e].sum(0), 2 * rows[b:e].sum(0), e - b)
    return t, p


def test_batch_means_allreduce():
    outs = _run(_means_job)
    g = torch.Generator().manual_seed(9)
    rows = torch.randn(10, 3, 512, generator=g)
    for t, p in outs:
        assert float((t - rows.mean(0)).abs().max()) < 1e-6
        assert float((p - 2 * rows.mean(0)).abs().max()) < 1e-6
