"""Host logic of team_b200.exemplars.construct_exemplar (drop-in for BaseLearner._construct_exemplar, models/base.py:274-343)
on CPU: the GPU pick (`ops.herding_select`) is replaced by the numpy restatement of the reference's selection, everything
else - dataset access, feature extraction loop, grouping of the new classes into one call, memory bookkeeping, class means -
is the product code.  The resulting exemplar memory must equal the REAL reference's (tests/golden/herding.npz)."""
import types

import numpy as np
import torch

from oracle import learner_harness
from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs


def test_construct_exemplar_bookkeeping(golden, monkeypatch):
    from team_b200 import exemplars, ops
    case, g = CASES["herding"], golden("herding")
    data = case_inputs(case)["data"]
    calls = []

    def fake_select(features, m, group_sizes=None):
        sizes = [features.shape[0]] if group_sizes is None else list(group_sizes)
        calls.append(sizes)
        idx, em, cm, r0 = [], [], [], 0
        for n in sizes:
            p, mean, cmean = O.herding_select(features[r0:r0 + n].numpy(), m)
            idx.append(torch.from_numpy(p)); em.append(torch.from_numpy(mean)); cm.append(torch.from_numpy(cmean))
            r0 += n
        return torch.stack(idx), torch.stack(em), torch.stack(cm)

    monkeypatch.setattr(ops, "herding_select", fake_select)
    nc, m = case["n_classes"], case["m"]
    net = types.SimpleNamespace(eval=lambda: None, extract_vector=lambda x: x)
    learner = types.SimpleNamespace(_network=net, _device=torch.device("cpu"), _known_classes=0, _total_classes=nc,
                                    _data_memory=np.array([]), _targets_memory=np.array([]), feature_dim=512,
                                    _class_means=np.zeros((nc, 512)))
    exemplars.construct_exemplar(learner, learner_harness.FakeDataManager(data), m)
    assert calls == [[case["n_train"]] * nc]                        # all new classes in ONE call
    mem = np.asarray(learner._data_memory, dtype=np.int64).reshape(nc, m, 2)
    first = np.array([np.where(data.y["train"] == c)[0][0] for c in range(nc)])
    assert np.array_equal(mem[:, :, 1] - first[:, None], g["picked"])
    assert np.array_equal(np.asarray(learner._targets_memory), np.repeat(np.arange(nc), m))
    assert np.allclose(learner._class_means, g["class_means"], rtol=0, atol=1e-6)
    # a second task appends behind the first one's memory
    learner._known_classes, learner._total_classes = nc, nc
    exemplars.construct_exemplar(learner, learner_harness.FakeDataManager(data), m)      # no new classes: nothing changes
    assert len(learner._targets_memory) == nc * m
