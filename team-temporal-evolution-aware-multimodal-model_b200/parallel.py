"""Data-parallel host logic of the path: one process per GPU, samples sharded
contiguously across ranks, weights / prototypes / prompts replicated.

Exactly two collectives exist on the path, both plain ``torch.distributed`` all-reduces
(NCCL over NVLink on the GPU box, gloo in the CPU tests):

  * prototype build: local keyed sums ``[K,512]`` fp32 + counts ``[K]`` int64 are summed
    over ranks, then divided (counts stay exact integers)      - models/proof.py:258-276
  * training step: the flat head-gradient bucket (``HeadStepRunner.flat_grads``,
    1.85 M fp32) is summed over ranks                            - models/proof.py:444

The reference has no working multi-GPU path (its ``nn.DataParallel`` wrap crashes,
models/proof.py:312-313 vs :248), so this module is new surface, not a mirror.
The graph / state-distance path is tiny (<= 60 nodes): every rank computes it redundantly
from the already all-reduced prototypes ("replicas only"), no collective.

Nothing here touches the CUDA library: the functions take the tensors the kernels
produced, so the same code runs under gloo on CPU tensors in ``tests/test_parallel_gloo.py``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_bounds(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``n_rows`` samples owned by ``rank``: the first
    ``n_rows % world_size`` ranks get one extra row (ragged batches keep every row exactly once)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(int(n_rows), world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_rows(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    b, e = shard_bounds(t.shape[0], rank, world_size)
    return t[b:e]


def allreduce_prototype_sums(sums: torch.Tensor, counts: torch.Tensor, group=None):
    """In-place sum over ranks of the local keyed sums (fp32) and counts (int64).
    Counts are integers, so the global per-class counts are exact whatever the reduction order."""
    if counts.dtype != torch.int64:
        raise TypeError("counts must be int64")
    _, ws = world(group)
    if ws > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return sums, counts


def allreduce_gradients(flat_grads: torch.Tensor, group=None, average: bool = False,
                        async_op: bool = False):
    """Sum (or mean) of the flat gradient bucket over ranks, in place.  The head's outputs are
    per-sample and every loss the learner uses is a batch mean, so callers that scale their
    cotangents by 1/global_batch want ``average=False`` (plain sum)."""
    _, ws = world(group)
    if ws == 1:
        return None
    work = dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    if average:
        if async_op:
            work.wait()
            work = None
        flat_grads.div_(ws)
    return work


def allreduce_batch_means(text_sum: torch.Tensor, proto_sum: torch.Tensor, local_rows: int, group=None):
    """PROOF ``Proof_Net.forward`` returns batch MEANS of the text / prototype rows
    (utils/inc_net.py:458-459); under data parallelism the exact single-GPU value needs the sums
    over all shards divided by the global batch.  Takes per-shard SUMS, returns global means."""
    n = torch.tensor([local_rows], dtype=torch.int64, device=text_sum.device)
    _, ws = world(group)
    if ws > 1:
        dist.all_reduce(text_sum, group=group)
        dist.all_reduce(proto_sum, group=group)
        dist.all_reduce(n, group=group)
    g = float(n.item())
    return text_sum / g, proto_sum / g
