"""Feature cache for the frozen CLIP towers.

The reference learner runs the image tower THREE times on the same batch (models/proof.py:418 via
forward_for_classification, :425 via forward_tri_modal, :430 via encode_image) and the text tower on all class
prompts plus twice on the per-sample prompts of every batch (:417, :425, :428) although a task has at most C distinct
prompts (utils/inc_net.py:403, :411 re-run the tower on every call).  The towers are frozen for the whole task
(models/proof.py:353-355), so their outputs are pure functions of the input:

  * ``TowerCache.image(x)``      one tower pass per distinct input tensor: a call with the tensor object of the previous call
                                 (same storage, same version counter) returns the stored features;
  * ``TowerCache.text(tokens)``  the tower only ever sees token rows it has not seen before (rows are matched exactly,
                                 on the device); a batch is assembled by a row gather;
  * ``IndexedFeatureStore``      image features by dataset index for loops that hand the index over (the loader already
                                 yields it, models/proof.py:403-411 drops it): one tower pass per sample per TASK.

Everything is keyed on the towers' parameter version counters as well: an optimiser step on (or a load_state_dict into)
the tower drops the cache.  The cache is bypassed while any tower parameter other than logit_scale requires a gradient.
Host-side bookkeeping + gathers only - no arithmetic of the path happens here."""
from __future__ import annotations

from typing import Callable, Optional

import torch


def _tower_params(convnet) -> list:
    params = convnet.named_parameters() if hasattr(convnet, "named_parameters") else ()
    return [p for name, p in params if "logit_scale" not in name]


def _tower_state(params) -> tuple:
    """(frozen?, fingerprint of the parameter versions and storages)."""
    frozen, ver = True, 0
    for p in params:
        frozen = frozen and not p.requires_grad
        ver = (ver * 1000003 + p._version + 7 * p.data_ptr()) & 0xFFFFFFFFFFFF
    return frozen, ver


def _match_rows(rows: torch.Tensor, table: torch.Tensor, max_cells: int = 1 << 26) -> torch.Tensor:
    """Index of every row of ``rows`` [u, L] in ``table`` [U, L] (exact match), -1 where absent.  The [u, U, L] comparison
    is done in slabs of at most ``max_cells`` elements so that a full table never costs more than 64 MB of scratch."""
    u, L = rows.shape
    U = table.shape[0]
    slot = torch.full((u,), -1, dtype=torch.int64, device=rows.device)
    step = max(1, max_cells // max(1, U * L))
    for lo in range(0, u, step):
        eq = (rows[lo:lo + step].unsqueeze(1) == table.unsqueeze(0)).all(dim=2)     # [step, U]
        hit = eq.any(dim=1)
        slot[lo:lo + step] = torch.where(hit, eq.to(torch.int64).argmax(dim=1), slot[lo:lo + step])
    return slot


class TowerCache:
    def __init__(self, convnet, max_text_rows: int = 65536):
        self.convnet = convnet
        self._params = _tower_params(convnet)           # Parameter objects survive .to() / load_state_dict (their storage and version change)
        self.max_text_rows = max_text_rows
        self.stats = {"image_calls": 0, "image_tower_runs": 0, "text_rows": 0, "text_tower_rows": 0}
        self.clear()

    def clear(self):
        self._img_key, self._img_in, self._img_out = None, None, None
        self._tok_table: Optional[torch.Tensor] = None       # [U, L] token rows seen so far
        self._tok_feats: Optional[torch.Tensor] = None       # [U, D] their features
        self._ver = None

    def _usable(self) -> bool:
        frozen, ver = _tower_state(self._params)
        if ver != self._ver:
            self.clear()
            self._ver = ver
        return frozen

    # ------------------------------------------------------------------ image tower
    def image(self, x: torch.Tensor) -> torch.Tensor:
        self.stats["image_calls"] += 1
        if not self._usable():
            self.stats["image_tower_runs"] += 1
            return self.convnet.encode_image(x)
        key = (x.data_ptr(), x._version, tuple(x.shape), x.dtype, x.device)
        if key != self._img_key:
            with torch.no_grad():
                out = self.convnet.encode_image(x)
            self.stats["image_tower_runs"] += 1
            self._img_key, self._img_in, self._img_out = key, x, out      # holding x keeps its storage (the key) alive
        return self._img_out

    # ------------------------------------------------------------------ text tower
    def text(self, tokens) -> torch.Tensor:
        if not torch.is_tensor(tokens) or tokens.dim() != 2:
            return self.convnet.encode_text(tokens)
        n = tokens.shape[0]
        self.stats["text_rows"] += n
        if not self._usable() or n == 0:
            self.stats["text_tower_rows"] += n
            return self.convnet.encode_text(tokens)
        uniq, inverse = torch.unique(tokens, dim=0, return_inverse=True)           # [u, L], [n]
        if self._tok_table is not None and (self._tok_table.shape[1] != uniq.shape[1] or self._tok_table.dtype != uniq.dtype
                                            or self._tok_table.device != uniq.device):
            self._tok_table, self._tok_feats = None, None
        if self._tok_table is None:
            slot = torch.full((uniq.shape[0],), -1, dtype=torch.int64, device=uniq.device)
        else:
            slot = _match_rows(uniq, self._tok_table)
        new = (slot < 0).nonzero().flatten()
        if new.numel() > 0:
            with torch.no_grad():
                f = self.convnet.encode_text(uniq.index_select(0, new))
            self.stats["text_tower_rows"] += int(new.numel())
            base = 0 if self._tok_table is None else self._tok_table.shape[0]
            if base + new.numel() > self.max_text_rows:                              # bounded: start over with this batch
                self._tok_table, self._tok_feats, base = None, None, 0
                with torch.no_grad():
                    f = self.convnet.encode_text(uniq)
                slot = torch.arange(uniq.shape[0], device=uniq.device)
                self._tok_table, self._tok_feats = uniq, f
            else:
                self._tok_table = uniq.index_select(0, new) if self._tok_table is None else torch.cat([self._tok_table, uniq.index_select(0, new)])
                self._tok_feats = f if self._tok_feats is None else torch.cat([self._tok_feats, f])
                slot = slot.clone()
                slot[new] = base + torch.arange(new.numel(), device=uniq.device)
        return self._tok_feats.index_select(0, slot.index_select(0, inverse))


class IndexedFeatureStore:
    """Image-tower features by dataset index: ``get(idx, inputs)`` runs the tower only on the samples of the batch that are
    not stored yet.  One store per task (the exemplar memory changes the index space between tasks): ``clear()`` in
    ``after_task``.  Storage is fp32 [capacity, 512] on the device (4 M samples = 8 GB of the 180 GB)."""

    def __init__(self, encode: Callable[[torch.Tensor], torch.Tensor], capacity: int, device, dim: int = 512):
        self.encode = encode
        self.feats = torch.empty((capacity, dim), dtype=torch.float32, device=device)
        self.have = torch.zeros((capacity,), dtype=torch.bool, device=device)
        self.tower_rows = 0

    def clear(self):
        self.have.zero_()

    def get(self, idx: torch.Tensor, inputs: torch.Tensor) -> torch.Tensor:
        idx = idx.to(self.feats.device, dtype=torch.int64)
        miss = (~self.have.index_select(0, idx)).nonzero().flatten()
        if miss.numel() > 0:
            with torch.no_grad():
                f = self.encode(inputs.index_select(0, miss.to(inputs.device)).to(self.feats.device)).float()
            self.tower_rows += int(miss.numel())
            rows = idx.index_select(0, miss)
            self.feats.index_copy_(0, rows, f)
            self.have.index_fill_(0, rows, True)
        return self.feats.index_select(0, idx)
