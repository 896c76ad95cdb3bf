"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/team_b200.h declares (no compute calls - there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "team_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(team_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from team_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 8
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/team_b200.h but not exported: {missing}"
    assert lib.team_version() >= 100


def test_no_cpu_fallback():
    import torch
    from team_b200 import capi, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.TeamB200Error):
        ops.cosine_logits(torch.zeros(4, 512), torch.zeros(2, 512))
    with pytest.raises(capi.TeamB200Error):
        ops.keyed_sums(torch.zeros(4, 512), torch.zeros(4, dtype=torch.int64), num_classes=2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "team-temporal-evolution-aware-multimodal-model_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f
