"""CPU check of the Gram formulation of the table-query rows (tests/gram_table_rows_model.py) against the direct
per-(sample, row) arithmetic, in fp64: forward outputs, the three weighted dY sums and the four score-gradient dots."""
import pytest
import torch

import gram_table_rows_model as G


@pytest.mark.parametrize("B,C,D,seed", [(6, 4, 64, 0), (3, 20, 512, 1), (9, 1, 32, 2)])
def test_gram_equals_direct(B, C, D, seed):
    case = G.make_case(B, C, D, seed)
    a, b = G.direct(case), G.gram(case)
    for k in a:
        err = float((a[k] - b[k]).norm() / a[k].norm().clamp_min(1e-30))
        assert err < 1e-9, (k, err)
