"""Host side of the life-stage graph path: graph construction, temporal GCN, pairwise state
distances, prototype sync and the state-distance matrix, over the C ABI
(``team_tgcn_forward`` / ``team_pairwise_state_dist`` / ``team_sync_prototypes`` / ``team_dist_*``).

Mirrors, with the same argument meaning and return structure,
  InsectLifecycleModel.evolve_and_update        models/state_evolution.py:239-367
  InsectLifecycleModel._detect_evolution_type   models/state_evolution.py:53-66
  TemporalStateGCN.forward                      models/dynamic_modal_graph.py:239-266
  Proof_Net.evolve_state_prototypes/_sync_class_prototypes   utils/inc_net.py:582-617
  AdaptiveStateDistanceMatrix.get_distance_matrix/forward    utils/state_distance.py:65-144
  Learner.update_state_distance_matrix (EMA)    models/proof.py:666-675
  DynamicGCN.forward (eval)                     models/dynamic_modal_graph.py:131-163
including the reference's quirks: the shallow-copy aliasing that mutates the
caller's prototype dict, zero-weight intra-class edges that still count in the mean, the
double EMA visit, state id 0 excluded from the batch centres.

The node/edge enumeration is vectorised numpy on the host (the reference does it with O(N^2)
Python loops); everything numeric runs in the CUDA library.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import capi

NUM_STATES = capi.NUM_STATES
_LIFECYCLE_NAMES = ("larvae_to_adult", "nymph_to_adult", "adult_only", "unknown")


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def detect_evolution_type(state_ids: Sequence[int]) -> str:
    """models/state_evolution.py:53-66."""
    if 1 in state_ids and 4 in state_ids:
        return "larvae_to_adult"
    elif 3 in state_ids and 4 in state_ids:
        return "nymph_to_adult"
    elif 1 in state_ids:
        return "larvae_to_adult"
    elif 3 in state_ids:
        return "nymph_to_adult"
    elif 4 in state_ids:
        return "adult_only"
    return "unknown"


class EvolutionGraph:
    """Nodes in the reference's order (class dict order x that class's state insertion order) and the
    edges as a destination-sorted CSR whose per-destination order equals the reference's edge-list
    order (intra-class edges by source index, then inter-class edges by source index)."""

    def __init__(self, node_class, node_state, node_time, rowptr, src, weight, class_order):
        self.node_class: np.ndarray = node_class      # int64 [N]
        self.node_state: np.ndarray = node_state      # int64 [N]
        self.node_time: np.ndarray = node_time        # float64 [N] (python-float arithmetic of the reference)
        self.rowptr: np.ndarray = rowptr              # int32 [N+1]
        self.src: np.ndarray = src                    # int32 [E]
        self.weight: np.ndarray = weight              # float32 [E]
        self.class_order: List[int] = class_order     # classes with >= 2 states, in dict order

    @property
    def n_nodes(self) -> int:
        return int(self.node_class.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.src.shape[0])

    def edge_list(self) -> Tuple[np.ndarray, np.ndarray]:
        """(edge_index [2,E] int64, weights [E] float32) in the reference's edge-list order
        (models/state_evolution.py:296-316): intra-class (i,j) row-major, then inter-class row-major."""
        dst = np.repeat(np.arange(self.n_nodes, dtype=np.int64), np.diff(self.rowptr))
        src = self.src.astype(np.int64)
        inter = self.node_class[src] != self.node_class[dst]
        order = np.lexsort((dst, src, inter))
        return np.stack([src[order], dst[order]]), self.weight[order]


def build_evolution_graph(by_state: Dict[int, Dict[int, torch.Tensor]],
                          lifecycle_types: Optional[Dict[int, str]] = None) -> EvolutionGraph:
    """Node / edge enumeration of evolve_and_update (models/state_evolution.py:260-316), SURVEY App. A-7.
    ``lifecycle_types`` is updated in place like ``self.class_lifecycle_types``."""
    if lifecycle_types is None:
        lifecycle_types = {}
    ncls: List[int] = []
    nst: List[int] = []
    ntime: List[float] = []
    order: List[int] = []
    for c, sd in by_state.items():
        if len(sd) < 2:
            continue
        sids = sorted(sd.keys())
        lifecycle_types[c] = detect_evolution_type(sids)
        s2t = {s: i / max(1, len(sids) - 1) for i, s in enumerate(sids)}
        order.append(c)
        for s in sd.keys():
            ncls.append(int(c)); nst.append(int(s)); ntime.append(s2t[s])
    n = len(ncls)
    cls = np.asarray(ncls, dtype=np.int64)
    st = np.asarray(nst, dtype=np.int64)
    tm = np.asarray(ntime, dtype=np.float64)
    if n == 0:
        return EvolutionGraph(cls, st, tm, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), order)
    # ---- intra-class edges i -> j, time_i < time_j, w = 1 - |dt|  (classes are contiguous node ranges)
    starts = np.flatnonzero(np.r_[True, cls[1:] != cls[:-1]])
    sizes = np.diff(np.r_[starts, n])
    first = np.repeat(starts, sizes)                   # first node of each node's class
    size_of = np.repeat(sizes, sizes)
    idx = np.arange(n)
    e_src, e_dst, e_w, e_kind = [], [], [], []
    for off in range(1, int(sizes.max())):
        for sgn in (1, -1):                            # partner = node at +/- off inside the class range
            j = idx + sgn * off
            ok = (j >= first) & (j < first + size_of)
            i_ok, j_ok = idx[ok], j[ok]
            keep = tm[i_ok] < tm[j_ok]
            i_ok, j_ok = i_ok[keep], j_ok[keep]
            e_src.append(i_ok); e_dst.append(j_ok)
            e_w.append(1.0 - np.abs(tm[i_ok] - tm[j_ok]))
            e_kind.append(np.zeros(i_ok.shape[0], np.int8))
    # ---- inter-class edges: same state id, same lifecycle type, different class, w = 0.5
    lt_id = {name: k for k, name in enumerate(_LIFECYCLE_NAMES)}
    lt = np.asarray([lt_id.get(lifecycle_types.get(c), len(lt_id)) for c in ncls], dtype=np.int64)
    key = st * 16 + lt
    perm = np.argsort(key, kind="stable")
    ks = key[perm]
    gstart = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    gsize = np.diff(np.r_[gstart, n])
    for g0, k in zip(gstart, gsize):
        if k < 2:
            continue
        m = perm[g0:g0 + k]                            # ascending node ids (stable sort)
        S = np.tile(m, (k, 1))                         # row = destination, columns = candidate sources
        Dd = np.repeat(m, k).reshape(k, k)
        ok = cls[S] != cls[Dd]
        e_src.append(S[ok]); e_dst.append(Dd[ok])
        e_w.append(np.full(int(ok.sum()), 0.5)); e_kind.append(np.ones(int(ok.sum()), np.int8))
    if e_src:
        src = np.concatenate(e_src); dst = np.concatenate(e_dst)
        w = np.concatenate(e_w); kind = np.concatenate(e_kind)
    else:
        src = dst = np.zeros(0, np.int64); w = np.zeros(0); kind = np.zeros(0, np.int8)
    o = np.lexsort((src, kind, dst))
    src, dst, w = src[o], dst[o], w[o]
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=n), out=rowptr[1:])
    return EvolutionGraph(cls, st, tm, rowptr, src.astype(np.int32), w.astype(np.float32), order)


def graph_from_edge_list(n_nodes: int, edge_index, edge_weights, time_steps) -> EvolutionGraph:
    """EvolutionGraph from the reference's COO inputs of TemporalStateGCN.forward
    (edge_index [2,E] src/dst, edge_weights [E], time_steps [N,1]): destination-sorted, stable, so the
    per-destination accumulation order equals the reference's edge-loop order."""
    ei = np.asarray(edge_index.detach().cpu().numpy() if torch.is_tensor(edge_index) else edge_index, dtype=np.int64).reshape(2, -1)
    w = np.asarray(edge_weights.detach().cpu().numpy() if torch.is_tensor(edge_weights) else edge_weights, dtype=np.float32).reshape(-1)
    t = np.asarray(time_steps.detach().cpu().numpy() if torch.is_tensor(time_steps) else time_steps, dtype=np.float64).reshape(-1)
    if t.shape[0] != n_nodes or w.shape[0] != ei.shape[1]:
        raise ValueError("time_steps must have one entry per node and edge_weights one per edge")
    if ei.size and (ei.min() < 0 or ei.max() >= n_nodes):
        raise ValueError("edge_index out of range")
    o = np.argsort(ei[1], kind="stable")
    rowptr = np.zeros(n_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(ei[1], minlength=n_nodes), out=rowptr[1:])
    z = np.zeros(n_nodes, dtype=np.int64)
    return EvolutionGraph(z, z.copy(), t, rowptr, ei[0][o].astype(np.int32), w[o], [])


# --------------------------------------------------------------------------- temporal GCN
def _tgcn_weights(p: Dict[str, torch.Tensor], prefix: str, dev):
    """ctypes view of the TemporalStateGCN parameters (state_dict names)."""
    keep = []

    def f(name):
        t = p[prefix + "." + name].detach()
        if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=dev, dtype=torch.float32).contiguous()
        keep.append(t)
        return t.data_ptr()

    w = capi.TgcnWeights()
    w.node_w, w.node_b = f("node_encoder.0.weight"), f("node_encoder.0.bias")
    w.node_ln_g, w.node_ln_b = f("node_encoder.1.weight"), f("node_encoder.1.bias")
    w.time_w, w.time_b = f("time_encoder.0.weight"), f("time_encoder.0.bias")
    w.time_ln_g, w.time_ln_b = f("time_encoder.1.weight"), f("time_encoder.1.bias")
    nb = 0
    while f"{prefix}.temporal_blocks.{nb}.message_net.0.weight" in p:
        if nb >= 8:
            raise ValueError("TemporalStateGCN: at most 8 blocks supported")
        b, q = w.blocks[nb], f"temporal_blocks.{nb}."
        b.msg_w, b.msg_b = f(q + "message_net.0.weight"), f(q + "message_net.0.bias")
        b.msg_ln_g, b.msg_ln_b = f(q + "message_net.1.weight"), f(q + "message_net.1.bias")
        b.upd_w, b.upd_b = f(q + "update_net.0.weight"), f(q + "update_net.0.bias")
        b.upd_ln_g, b.upd_ln_b = f(q + "update_net.1.weight"), f(q + "update_net.1.bias")
        b.gate_w, b.gate_b = f(q + "temporal_gate.0.weight"), f(q + "temporal_gate.0.bias")
        nb += 1
    w.num_blocks = nb
    w.out_w, w.out_b = f("output_proj.weight"), f("output_proj.bias")
    return w, keep


def temporal_state_gcn(p: Dict[str, torch.Tensor], node_features: torch.Tensor, graph: EvolutionGraph,
                       prefix: str = "state_embedder.temporal_gcn") -> torch.Tensor:
    """TemporalStateGCN.forward (models/dynamic_modal_graph.py:239-266) -> [N,512] unit rows."""
    capi.require_device()
    if not node_features.is_cuda:
        raise capi.TeamB200Error("temporal_state_gcn needs CUDA tensors (no CPU fallback)")
    dev = node_features.device
    x = node_features.detach().to(torch.float32).contiguous()
    n = x.shape[0]
    if n != graph.n_nodes or x.shape[1] != capi.D:
        raise ValueError("node_features must be [n_nodes,512]")
    w, keep = _tgcn_weights(p, prefix, dev)
    tsteps = torch.from_numpy(graph.node_time.astype(np.float32)).to(dev)
    rowptr = torch.from_numpy(graph.rowptr).to(dev)
    src = torch.from_numpy(graph.src if graph.n_edges else np.zeros(1, np.int32)).to(dev)
    ew = torch.from_numpy(graph.weight if graph.n_edges else np.zeros(1, np.float32)).to(dev)
    L = capi.lib()
    nbytes = L.team_tgcn_workspace_bytes(n)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    out = torch.empty((n, capi.D), dtype=torch.float32, device=dev)
    capi.check(L.team_tgcn_forward(C.byref(w), x.data_ptr(), tsteps.data_ptr(), n, rowptr.data_ptr(), src.data_ptr(),
                                   ew.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, _stream_ptr()),
               "team_tgcn_forward")
    del keep
    return out


def _group_mean(nodes: torch.Tensor, group_ptr: np.ndarray, member: Optional[np.ndarray]) -> torch.Tensor:
    dev = nodes.device
    ng = group_ptr.shape[0] - 1
    out = torch.empty((ng, capi.D), dtype=torch.float32, device=dev)
    gp = torch.from_numpy(group_ptr.astype(np.int32)).to(dev)
    mb = torch.from_numpy(member.astype(np.int32)).to(dev) if member is not None else None
    capi.check(capi.lib().team_group_mean(nodes.data_ptr(), gp.data_ptr(), mb.data_ptr() if mb is not None else None,
                                          ng, out.data_ptr(), _stream_ptr()), "team_group_mean")
    return out


def pairwise_state_distances(nodes: torch.Tensor, node_state: np.ndarray) -> Dict[int, Dict[int, float]]:
    """mean over ordered node pairs i != j of 1 - cos(u_i,u_j), keyed [state_i][state_j]
    (models/state_evolution.py:345-364); ONE device->host copy instead of one per pair."""
    dev = nodes.device
    n = nodes.shape[0]
    st = torch.from_numpy(node_state.astype(np.int32)).to(dev)
    sums = torch.empty((NUM_STATES * NUM_STATES,), dtype=torch.float64, device=dev)
    counts = torch.empty((NUM_STATES * NUM_STATES,), dtype=torch.int64, device=dev)
    nbytes = n * NUM_STATES * 12 + 256
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    capi.check(capi.lib().team_pairwise_state_dist(nodes.data_ptr(), st.data_ptr(), n, sums.data_ptr(),
                                                   counts.data_ptr(), ws.data_ptr(), nbytes, _stream_ptr()),
               "team_pairwise_state_dist")
    s = sums.cpu().numpy().reshape(NUM_STATES, NUM_STATES)
    c = counts.cpu().numpy().reshape(NUM_STATES, NUM_STATES)
    # key order of the reference's nested dict: s1 by first node occurrence; the inner keys are fixed by the
    # first node i0 of state s1 (its j loop sees every other node): first-occurrence order over j != i0
    dist: Dict[int, Dict[int, float]] = {}
    states = [int(v) for v in node_state]
    for i0, s1 in enumerate(states):
        if s1 in dist:
            continue
        inner: Dict[int, float] = {}
        for j, s2 in enumerate(states):
            if j != i0 and s2 not in inner:
                inner[s2] = float(s[s1, s2] / c[s1, s2])
        dist[s1] = inner
    return dist


def evolve_and_update(p: Dict[str, torch.Tensor], by_state: Dict[int, Dict[int, torch.Tensor]],
                      lifecycle_types: Optional[Dict[int, str]] = None, epoch=None, max_epoch=None,
                      prefix: str = "state_embedder.temporal_gcn") -> dict:
    """InsectLifecycleModel.evolve_and_update (models/state_evolution.py:239-367).  The result's
    'prototypes' is a shallow copy of ``by_state``: the inner dicts ARE the caller's and receive the
    evolved rows."""
    capi.require_device()
    result = {"prototypes": by_state.copy(), "embeddings": [], "lifecycle_features": {}, "distances": {}}
    if len(by_state) < 1:
        return result
    if lifecycle_types is None:
        lifecycle_types = {}
    graph = build_evolution_graph(by_state, lifecycle_types)
    if graph.n_nodes == 0:
        return result
    rows = [by_state[c][s] for c in graph.class_order for s in by_state[c].keys()]
    dev = rows[0].device
    if not rows[0].is_cuda:
        raise capi.TeamB200Error("evolve_and_update needs CUDA prototypes (no CPU fallback)")
    x = torch.stack([r.detach().to(torch.float32) for r in rows]).contiguous()
    # lifecycle features: mean over each class's states in ascending state order (:256-258)
    starts = np.flatnonzero(np.r_[True, graph.node_class[1:] != graph.node_class[:-1]])
    gptr = np.r_[starts, graph.n_nodes].astype(np.int32)
    member = np.concatenate([starts[k] + np.argsort(graph.node_state[gptr[k]:gptr[k + 1]], kind="stable")
                             for k in range(len(starts))])
    lf = _group_mean(x, gptr, member)
    for k, c in enumerate(graph.class_order):
        result["lifecycle_features"][c] = lf[k]
    if graph.n_edges == 0:
        return result
    upd = temporal_state_gcn(p, x, graph, prefix)
    for i, (c, s) in enumerate(zip(graph.node_class.tolist(), graph.node_state.tolist())):
        result["prototypes"][c][s] = upd[i]
    emb = _group_mean(upd, gptr, None)
    for k, c in enumerate(graph.class_order):
        while len(result["embeddings"]) <= c:
            result["embeddings"].append(None)
        result["embeddings"][c] = emb[k]
    result["distances"] = pairwise_state_distances(upd, graph.node_state)
    return result


def sync_class_prototypes(img_prototypes: torch.Tensor, by_state: Dict[int, Dict[int, torch.Tensor]]):
    """Proof_Net._sync_class_prototypes (utils/inc_net.py:600-617), in place on ``img_prototypes``."""
    capi.require_device()
    rows, states, gptr, gcls = [], [], [0], []
    for c in range(len(img_prototypes)):
        if c in by_state and by_state[c]:
            for s, pr in by_state[c].items():
                rows.append(pr); states.append(int(s))
            gptr.append(len(rows)); gcls.append(c)
    if not rows:
        return img_prototypes
    if not img_prototypes.is_cuda or img_prototypes.dtype != torch.float32 or not img_prototypes.is_contiguous():
        raise capi.TeamB200Error("img_prototypes must be a contiguous fp32 CUDA tensor")
    dev = img_prototypes.device
    x = torch.stack([r.detach().to(device=dev, dtype=torch.float32) for r in rows]).contiguous()
    t = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
    gp, ns, gc = t(gptr), t(states), t(gcls)
    capi.check(capi.lib().team_sync_prototypes(x.data_ptr(), gp.data_ptr(), ns.data_ptr(), gc.data_ptr(), len(gcls),
                                               img_prototypes.data_ptr(), _stream_ptr()), "team_sync_prototypes")
    return img_prototypes


def evolve_state_prototypes(p: Dict[str, torch.Tensor], img_prototypes: torch.Tensor,
                            by_state: Dict[int, Dict[int, torch.Tensor]],
                            lifecycle_types: Optional[Dict[int, str]] = None):
    """Proof_Net.evolve_state_prototypes (utils/inc_net.py:582-598).  Because of the aliasing above,
    ``0.6*original + 0.4*evolved`` is ``evolved``; what remains is the re-normalisation, done on the
    device for every row the GCN wrote, then the class-prototype sync."""
    if not by_state:
        return None
    res = evolve_and_update(p, by_state, lifecycle_types)
    L = capi.lib()
    # The rows the GCN wrote are views of ONE [nodes, 512] buffer (evolve_and_update): normalise that buffer with one launch
    # instead of one launch per (class, state) row - 466 launches at 200 classes; rows that are tensors of their own
    # (classes with a single state are not part of the graph) keep a launch each.
    whole: Dict[int, list] = {}
    for c, sp in res["prototypes"].items():
        for s, ev in sp.items():
            if c in by_state and s in by_state[c]:
                row = by_state[c][s]
                if not (row.is_cuda and row.dtype == torch.float32 and row.is_contiguous()):
                    raise capi.TeamB200Error("state prototypes must be contiguous fp32 CUDA rows")
                base = row._base
                if (base is not None and base.dim() == 2 and base.shape[1] == row.shape[0] and base.is_contiguous()
                        and base.dtype == torch.float32):
                    whole.setdefault(id(base), [base, 0])[1] += 1
                else:
                    capi.check(L.team_rows_normalize(row.data_ptr(), 1, _stream_ptr()), "team_rows_normalize")
    for base, n in whole.values():
        if n == base.shape[0]:                 # every row of the buffer is a prototype: one launch
            capi.check(L.team_rows_normalize(base.data_ptr(), base.shape[0], _stream_ptr()), "team_rows_normalize")
        else:                                  # a buffer only partly referenced: row by row, touching nothing else
            for c, sp in res["prototypes"].items():
                for s, ev in sp.items():
                    row = by_state.get(c, {}).get(s)
                    if row is not None and row._base is base:
                        capi.check(L.team_rows_normalize(row.data_ptr(), 1, _stream_ptr()), "team_rows_normalize")
    sync_class_prototypes(img_prototypes, by_state)
    return res["embeddings"]


# --------------------------------------------------------------------------- state-distance matrix
def prior_distance_factors(num_states: int = NUM_STATES, device=None) -> torch.Tensor:
    """AdaptiveStateDistanceMatrix prior (utils/state_distance.py:20-37) - host constants."""
    m = np.ones((num_states, num_states), dtype=np.float32)
    m[1, 4] = m[4, 1] = 2.0
    m[3, 4] = m[4, 3] = 0.7
    m[1, 2] = m[2, 1] = 1.5
    m[0, :] = 1.8
    m[:, 0] = 1.8
    m[0, 0] = 1.0
    return torch.from_numpy(m).to(device) if device is not None else torch.from_numpy(m)


def get_distance_matrix(factors: torch.Tensor) -> torch.Tensor:
    """utils/state_distance.py:65-71: (F + F^T)/2 with unit diagonal."""
    capi.require_device()
    f = factors.detach()
    if not f.is_cuda or f.dtype != torch.float32 or not f.is_contiguous():
        raise capi.TeamB200Error("distance_factors must be a contiguous fp32 CUDA tensor")
    out = torch.empty_like(f)
    capi.check(capi.lib().team_dist_matrix(f.data_ptr(), f.shape[0], out.data_ptr(), _stream_ptr()), "team_dist_matrix")
    return out


def update_state_distance_matrix(factors: torch.Tensor, distances: Dict[int, Dict[int, float]], weight: float = 0.3):
    """The Learner's EMA (models/proof.py:666-675), sequential in dict order, in place on ``factors``
    (each unordered pair is visited twice)."""
    capi.require_device()
    keys, vals = [], []
    for s1 in distances:
        for s2 in distances[s1]:
            keys += [int(s1), int(s2)]; vals.append(float(distances[s1][s2]))
    if not vals:
        return factors
    dev = factors.device
    k = torch.tensor(keys, dtype=torch.int32, device=dev)
    v = torch.tensor(vals, dtype=torch.float64, device=dev)
    capi.check(capi.lib().team_dist_ema(factors.data_ptr(), factors.shape[0], k.data_ptr(), v.data_ptr(), len(vals),
                                        float(weight), _stream_ptr()), "team_dist_ema")
    return factors


def state_distance_forward(factors: torch.Tensor, state_features: torch.Tensor, state_ids: torch.Tensor,
                           update_counter: int, training: bool = True, update_interval: int = 10,
                           decay: float = 0.9):
    """AdaptiveStateDistanceMatrix.forward (utils/state_distance.py:79-144): returns the PRE-update
    matrix and the new counter; every ``update_interval``-th training call updates ``factors`` in place
    from the batch's per-state centres (states 1..9)."""
    from . import ops
    cur = get_distance_matrix(factors)
    if training and update_counter % update_interval == 0:
        sums, counts = ops.keyed_sums(state_features.detach(), state_ids, num_classes=factors.shape[0])
        capi.check(capi.lib().team_state_dist_forward(sums.data_ptr(), counts.data_ptr(), factors.data_ptr(),
                                                      float(decay), None, _stream_ptr()), "team_state_dist_forward")
    return cur, update_counter + 1


# --------------------------------------------------------------------------- DynamicGCN
def dynamic_gcn(x: torch.Tensor, edge_index: Optional[torch.Tensor], edge_weights: Optional[torch.Tensor],
                layers: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """DynamicGCN.forward in eval mode (models/dynamic_modal_graph.py:131-163).
    ``layers`` = [(weight, bias, ln_weight, ln_bias), ...]; per layer h = relu(Linear(x)),
    h[dst] += w * h_pre[src] in edge order, LayerNorm."""
    capi.require_device()
    if not x.is_cuda:
        raise capi.TeamB200Error("dynamic_gcn needs CUDA tensors (no CPU fallback)")
    dev = x.device
    n = x.shape[0]
    xs = x.detach().to(torch.float32).contiguous()
    keep = []
    arr = (capi.DgcnLayer * len(layers))()
    maxd = 0
    for i, (W, b, g, be) in enumerate(layers):
        ts = [t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in (W, b, g, be)]
        keep += ts
        arr[i].w, arr[i].b, arr[i].ln_g, arr[i].ln_b = (t.data_ptr() for t in ts)
        arr[i].in_dim, arr[i].out_dim = int(W.shape[1]), int(W.shape[0])
        maxd = max(maxd, int(W.shape[0]))
    rp = sr = ew = None
    if edge_index is not None and edge_weights is not None:
        ei = edge_index.detach().cpu().numpy().astype(np.int64)
        w = edge_weights.detach().cpu().numpy().astype(np.float32)
        o = np.argsort(ei[1], kind="stable")                      # dst-sorted, edge order kept inside a destination
        rowptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(ei[1], minlength=n), out=rowptr[1:])
        rp = torch.from_numpy(rowptr).to(dev)
        sr = torch.from_numpy(ei[0][o].astype(np.int32) if o.size else np.zeros(1, np.int32)).to(dev)
        ew = torch.from_numpy(w[o] if o.size else np.zeros(1, np.float32)).to(dev)
    L = capi.lib()
    nbytes = L.team_dgcn_workspace_bytes(n, maxd)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    out = torch.empty((n, int(layers[-1][0].shape[0])), dtype=torch.float32, device=dev)
    capi.check(L.team_dgcn_forward(arr, len(layers), xs.data_ptr(), n, rp.data_ptr() if rp is not None else None,
                                   sr.data_ptr() if sr is not None else None, ew.data_ptr() if ew is not None else None,
                                   out.data_ptr(), ws.data_ptr(), nbytes, _stream_ptr()), "team_dgcn_forward")
    del keep
    return out
