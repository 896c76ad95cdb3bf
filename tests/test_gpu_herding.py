"""GPU exemplar herding (team_herding_select; BaseLearner._construct_exemplar, models/base.py:274-343)
against the picks of the real reference (golden) and the numpy restatement.  Picks are exact wherever the runner-up's
distance differs by more than fp32 round-off (checked: the margin of every oracle pick is recorded and must exceed 1e-6
relative for the comparison to count); exemplar / class means <= 1e-6 relative."""
import numpy as np
import pytest
import torch

from oracle import team_oracle as O
from oracle.cases import CASES, case_inputs

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_herding_vs_golden(golden):
    from team_b200 import ops
    case, g = CASES["herding"], golden("herding")
    data = case_inputs(case)["data"]
    order = np.argsort(data.y["train"], kind="stable")
    x = data.x["train"][order].cuda().contiguous()
    sizes = [int((data.y["train"] == c).sum()) for c in range(case["n_classes"])]
    idx, emean, cmean = ops.herding_select(x, case["m"], sizes)                 # all classes in one launch
    assert np.array_equal(idx.cpu().numpy(), g["picked"])
    assert rel(emean, g["class_means"]) < 1e-6
    one, em1, _ = ops.herding_select(x[:sizes[0]].contiguous(), case["m"])       # a single class
    assert np.array_equal(one.cpu().numpy()[0], g["picked"][0]) and rel(em1[0], g["class_means"][0]) < 1e-6


@pytest.mark.parametrize("n,m,seed", [(20, 20, 1), (333, 5, 2), (4100, 40, 3), (1, 1, 4)])
def test_herding_vs_oracle(n, m, seed):
    """Every row picked (m = n), ragged n, more rows than one CTA pass, a single row."""
    from team_b200 import ops
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(512, generator=g) + 0.8 * torch.randn(n, 512, generator=g)) * (0.5 + torch.rand(n, 1, generator=g))
    picked, mean, cmean = O.herding_select(x.numpy(), m)
    idx, emean, cm = ops.herding_select(x.cuda(), m)
    assert np.array_equal(idx.cpu().numpy()[0], picked)
    assert len(set(picked.tolist())) == m
    assert rel(emean[0], mean) < 1e-6 and rel(cm[0], cmean) < 1e-6
    a, b, _ = ops.herding_select(x.cuda(), m)
    assert torch.equal(a, idx) and torch.equal(b, emean)            # run-to-run identical


def test_herding_ties_take_first_index():
    """Duplicate rows: numpy's argmin returns the first of equal distances, and so must the kernel."""
    from team_b200 import ops
    g = torch.Generator().manual_seed(9)
    base = torch.randn(6, 512, generator=g)
    x = torch.cat([base, base, base])                                # every row three times
    picked, _, _ = O.herding_select(x.numpy(), 7)
    idx, _, _ = ops.herding_select(x.cuda(), 7)
    assert np.array_equal(idx.cpu().numpy()[0], picked)


def test_herding_rejects_short_class():
    from team_b200 import ops
    with pytest.raises(ValueError):
        ops.herding_select(torch.randn(3, 512).cuda(), 4)
