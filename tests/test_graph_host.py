"""CPU checks of the host-side graph construction (team_b200.graph.build_evolution_graph) against the
oracle's enumeration and the reference-generated golden edge lists (models/state_evolution.py:260-316)."""
import numpy as np
import pytest

from oracle import synth
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs


@pytest.mark.parametrize("name", ["evolve_6cls", "evolve_20cls"])
def test_graph_matches_reference_golden(name, golden):
    from team_b200 import graph
    ci, g = case_inputs(CASES[name]), golden(name)
    lt = {}
    eg = graph.build_evolution_graph(ci["by_state"], lt)
    ei, ew = eg.edge_list()
    assert np.array_equal(ei, g["edge_index"])
    assert np.array_equal(ew, g["edge_weights"])
    assert np.array_equal(eg.node_time.astype(np.float32).reshape(-1, 1), g["time_steps"])
    assert [tuple(k) for k in g["proto_keys"].tolist()][:eg.n_nodes] == list(zip(eg.node_class.tolist(), eg.node_state.tolist()))


@pytest.mark.parametrize("nc,pattern", [
    (3, ((1, 4), (3, 4), (1, 2, 4))),
    (37, ((4, 1), (3, 4), (1, 2, 4), (4,), (2, 5), (5, 3, 1, 4))),      # unsorted insertion order, 1-state and 'unknown' classes
    (5, ((4,),)),                                                      # no class with two states -> empty graph
])
def test_graph_matches_oracle(nc, pattern):
    from team_b200 import graph
    bs = synth.make_state_prototype_dict(nc, pattern=pattern)
    ncls, nst, ntime, edges, w, lt_ref = O.build_evolution_graph(bs)
    lt = {}
    eg = graph.build_evolution_graph(bs, lt)
    assert lt == lt_ref
    assert eg.node_class.tolist() == ncls and eg.node_state.tolist() == nst
    assert np.array_equal(eg.node_time, np.array(ntime, dtype=np.float64))
    ei, ew = eg.edge_list()
    assert ei.shape[1] == len(edges)
    if edges:
        assert np.array_equal(ei, np.array(edges, dtype=np.int64).T)
        assert np.array_equal(ew, np.array(w, dtype=np.float32))
    for d in range(eg.n_nodes):        # CSR keeps the reference edge order inside every destination
        assert eg.src[eg.rowptr[d]:eg.rowptr[d + 1]].tolist() == [e[0] for e in edges if e[1] == d]


def test_known_example():
    """SURVEY App. A-7 probed example: 3 classes {1,4},{3,4},{1,2,4} -> 7 nodes, 9 edges."""
    import torch
    from team_b200 import graph
    z = torch.zeros(512)
    eg = graph.build_evolution_graph({0: {1: z, 4: z}, 1: {3: z, 4: z}, 2: {1: z, 2: z, 4: z}})
    ei, ew = eg.edge_list()
    assert list(zip(ei[0].tolist(), ei[1].tolist(), ew.tolist())) == [
        (0, 1, 0.0), (2, 3, 0.0), (4, 5, .5), (4, 6, 0.0), (5, 6, .5), (0, 4, .5), (1, 6, .5), (4, 0, .5), (6, 1, .5)]
    assert graph.detect_evolution_type([2, 5]) == "unknown"
