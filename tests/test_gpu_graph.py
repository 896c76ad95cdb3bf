"""GPU parity of the life-stage graph path (temporal GCN, pairwise state distances, prototype sync,
state-distance matrix, DynamicGCN) through the C ABI against the reference-generated golden vectors
and the CPU oracle.  fp32: 1e-5 relative (norm-wise); keys, counts and dict orders exact."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _dev_state_dict(by_state, dev):
    return {c: {s: v.clone().to(dev) for s, v in sd.items()} for c, sd in by_state.items()}


@pytest.mark.parametrize("name", ["evolve_6cls", "evolve_20cls"])
def test_evolve_and_update_vs_golden(name, golden):
    from team_b200 import graph
    dev = torch.device("cuda")
    ci, g = case_inputs(CASES[name]), golden(name)
    p = {k: v.to(dev) for k, v in ci["params"].items()}
    by_state = _dev_state_dict(ci["by_state"], dev)
    lifecycle = {}
    res = graph.evolve_and_update(p, by_state, lifecycle)
    keys = [(c, s) for c, sd in res["prototypes"].items() for s in sd]
    assert keys == [tuple(k) for k in g["proto_keys"].tolist()]
    vals = torch.stack([res["prototypes"][c][s] for c, s in keys])
    assert rel(vals, g["proto_vals"]) < 1e-5, rel(vals, g["proto_vals"])
    assert all(by_state[c][s] is res["prototypes"][c][s] for c, s in keys)       # aliasing quirk, App. C-2
    emb_idx = [i for i, e in enumerate(res["embeddings"]) if e is not None]
    assert emb_idx == g["emb_idx"].tolist()
    assert rel(torch.stack([res["embeddings"][i] for i in emb_idx]), g["emb_vals"]) < 1e-5
    assert list(res["lifecycle_features"].keys()) == g["lifecycle_idx"].tolist()
    assert rel(torch.stack(list(res["lifecycle_features"].values())), g["lifecycle_vals"]) < 1e-5
    dk = [(s1, s2) for s1, dd in res["distances"].items() for s2 in dd]
    assert dk == [tuple(k) for k in g["dist_keys"].tolist()]
    dv = np.array([res["distances"][a][b] for a, b in dk])
    assert np.allclose(dv, g["dist_vals"], rtol=0, atol=1e-5)
    # second GCN pass + (no-op) blend + re-normalise + class-prototype sync
    protos = ci["protos"].clone().to(dev)
    graph.evolve_state_prototypes(p, protos, by_state, lifecycle)
    assert rel(protos, g["img_prototypes_after_sync"]) < 1e-5
    # prior + the Learner's double-visit EMA
    f = graph.prior_distance_factors(device=dev)
    assert np.array_equal(f.cpu().numpy(), g["prior_factors"])
    assert np.array_equal(graph.get_distance_matrix(f).cpu().numpy(), g["prior_matrix"])
    res3 = graph.evolve_and_update(p, by_state, lifecycle)
    graph.update_state_distance_matrix(f, res3["distances"])
    assert np.allclose(f.cpu().numpy(), g["factors_after_update"], rtol=0, atol=1e-5)
    assert np.allclose(graph.get_distance_matrix(f).cpu().numpy(), g["matrix_after_update"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("nc,pattern", [(37, ((4, 1), (3, 4), (1, 2, 4), (4,), (2, 5), (5, 3, 1, 4))),
                                        (400, ((1, 4), (3, 4), (1, 2, 4)))])
def test_evolve_vs_oracle(nc, pattern):
    """ragged / unsorted / unknown-lifecycle classes and a scaled graph (920 nodes) against the fp64 oracle GCN."""
    from team_b200 import graph
    dev = torch.device("cuda")
    params = synth.make_params(2, seed=77)
    bs = synth.make_state_prototype_dict(nc, seed=9, pattern=pattern)
    eg = graph.build_evolution_graph(bs, {})
    x = torch.stack([bs[c][s] for c in eg.class_order for s in bs[c].keys()])
    out = graph.temporal_state_gcn({k: v.to(dev) for k, v in params.items()}, x.to(dev), eg)
    ei, ew = eg.edge_list()
    if nc <= 64:       # the oracle's per-edge Python loop is only affordable for small graphs
        p64 = {k: v.double() for k, v in params.items()}
        ref = O.temporal_state_gcn(x.double(), torch.from_numpy(ei), torch.from_numpy(ew).double(),
                                   torch.from_numpy(eg.node_time).reshape(-1, 1), p64)
        assert rel(out, ref) < 1e-5, rel(out, ref)
        d = graph.pairwise_state_distances(out, eg.node_state)
        r = O.evolve_and_update({k: v.clone() for k, v in params.items()},
                                {c: {s: v.clone() for s, v in sd.items()} for c, sd in bs.items()}, {})["distances"]
        assert [(a, b) for a in d for b in d[a]] == [(a, b) for a in r for b in r[a]]
        assert np.allclose([d[a][b] for a in d for b in d[a]], [r[a][b] for a in r for b in r[a]], atol=1e-5)
    else:              # size-independent properties: unit rows, finite, deterministic
        assert torch.isfinite(out).all()
        assert torch.allclose(out.norm(dim=1), torch.ones(out.shape[0], device=dev), atol=1e-5)
        out2 = graph.temporal_state_gcn({k: v.to(dev) for k, v in params.items()}, x.to(dev), eg)
        assert torch.equal(out, out2)


def test_state_distance_forward_vs_golden(golden):
    from team_b200 import graph
    dev = torch.device("cuda")
    ci, g = case_inputs(CASES["state_distance_forward"]), golden("state_distance_forward")
    f = graph.prior_distance_factors(device=dev)
    ret0, cnt = graph.state_distance_forward(f, ci["feat"].to(dev), ci["sid"].to(dev), 0)
    assert np.array_equal(ret0.cpu().numpy(), g["ret0"])
    assert np.allclose(f.cpu().numpy(), g["factors1"], rtol=0, atol=2e-6)
    ret1, cnt = graph.state_distance_forward(f, ci["feat"].to(dev), ci["sid"].to(dev), cnt)
    assert np.allclose(ret1.cpu().numpy(), g["ret1"], rtol=0, atol=2e-6)
    assert np.allclose(f.cpu().numpy(), g["factors2"], rtol=0, atol=2e-6)
    assert cnt == int(g["counter"])
    assert float(f[0, 1]) == np.float32(1.8)


def test_dynamic_gcn_vs_golden(golden):
    from team_b200 import graph
    dev = torch.device("cuda")
    ci, g = case_inputs(CASES["dynamic_gcn"]), golden("dynamic_gcn")
    out = graph.dynamic_gcn(ci["x"].to(dev), ci["edge_index"], ci["edge_weights"], ci["layers"])
    assert rel(out, g["out"]) < 1e-5, rel(out, g["out"])
    out_noedge = graph.dynamic_gcn(ci["x"].to(dev), None, None, ci["layers"])
    assert rel(out_noedge, O.dynamic_gcn(ci["x"], None, None, ci["layers"])) < 1e-5


def test_sync_prototypes_vs_oracle():
    from team_b200 import graph
    dev = torch.device("cuda")
    bs = synth.make_state_prototype_dict(12, seed=5, pattern=((1, 4), (4,), (3, 2, 4), (2,)))
    del bs[7]
    protos = synth.make_prototypes(14, seed=11)
    ref = O.sync_class_prototypes(protos.clone(), bs)
    out = graph.sync_class_prototypes(protos.clone().to(dev), _dev_state_dict(bs, dev))
    assert rel(out, ref) < 1e-6
    assert torch.equal(out[7].cpu(), protos[7]) and torch.equal(out[13].cpu(), protos[13])     # untouched rows
