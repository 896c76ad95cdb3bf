"""Exemplar memory of the learner on the GPU.

``construct_exemplar`` is a drop-in for ``BaseLearner._construct_exemplar`` (models/base.py:274-343): same arguments, same
effect on ``_data_memory`` / ``_targets_memory`` / ``_class_means``.  The reference extracts every new class through the
network, moves the features to the host and runs an O(m n) numpy loop per class, then extracts the chosen exemplars a
second time for their mean; here the features of all new classes stay on the device and ONE launch of
``team_herding_select`` (one CTA per class) returns the picks and the exemplar means.  Bind it with

    models.base.BaseLearner._construct_exemplar = team_b200.exemplars.construct_exemplar
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import ops

BATCH_SIZE = 128          # models/base.py:11


def _features(learner, loader) -> torch.Tensor:
    """BaseLearner._extract_vectors (models/base.py:214-236) without the host round trip."""
    learner._network.eval()
    out = []
    with torch.no_grad():
        for _, inputs, _targets in loader:
            if isinstance(inputs, dict):
                inputs = inputs["image"]
            out.append(learner._network.extract_vector(inputs.to(learner._device)).float())
    return torch.cat(out)


def construct_exemplar(learner, data_manager, m: int, num_workers: int = 0):
    classes = list(range(learner._known_classes, learner._total_classes))
    if not classes:
        return
    data_of, feats = {}, []
    for c in classes:
        data, _targets, ds = data_manager.get_dataset(np.arange(c, c + 1), source="train", mode="test", ret_data=True)
        data_of[c] = data
        feats.append(_features(learner, DataLoader(ds, batch_size=BATCH_SIZE, shuffle=False, num_workers=num_workers)))
    sizes = [int(f.shape[0]) for f in feats]
    idx, emean, _ = ops.herding_select(torch.cat(feats).contiguous(), m, sizes)
    idx, emean = idx.cpu().numpy(), emean.double().cpu().numpy()
    for g, c in enumerate(classes):
        selected = np.array([np.array(data_of[c][i]) for i in idx[g]])
        targets = np.full(m, c)
        learner._data_memory = np.concatenate((learner._data_memory, selected)) if len(learner._data_memory) != 0 else selected
        learner._targets_memory = (np.concatenate((learner._targets_memory, targets))
                                   if len(learner._targets_memory) != 0 else targets)
        learner._class_means[c, :] = emean[g]
