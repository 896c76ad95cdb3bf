"""CPU checks of the drop-in boundary beyond symbol names: the ctypes mirrors in team_b200/capi.py must have exactly
the size and field offsets of the C structs in include/team_b200.h (a small C program compiled with gcc prints them),
the header must compile as plain C, and the flat gradient buffer must keep the bucket order the backward's folded
gradient exchange relies on (team_head_grads.comm: w_fc, then w_q / w_k / w_v, then everything else)."""
import ctypes
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PAIRS = [  # (C struct, ctypes class name, fields to compare: C field -> ctypes field)
    ("team_head_weights", "HeadWeights", ["num_tasks", "prompts_per_task", "w_img", "prompts", "state_emb", "w_q", "w_fc", "ln_b", "prototypes", "num_classes", "num_frozen", "w_frozen", "b_frozen"]),
    ("team_head_grads", "HeadGrads", ["w_img", "prompts", "state_emb", "w_q", "w_fc", "ln_b", "ev_w_fc", "ev_w_qkv", "comm", "g_own_rows"]),
    ("team_peer_comm", "PeerComm", ["bufs", "flags", "multicast", "rank", "world", "n_total", "split_at"]),
    ("team_tgcn_block", "TgcnBlock", ["msg_w", "gate_b"]),
    ("team_dgcn_layer", "DgcnLayer", []),
    ("team_gemm_desc", "GemmDesc", ["M", "N", "K", "A", "C", "bias"]),
    ("team_tgcn_weights", "TgcnWeights", []),
]


def test_ctypes_mirrors_match_the_c_header():
    from team_b200 import capi
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "team_b200.h"', "int main(void) {"]
    for cs, _, fields in PAIRS:
        lines.append(f'  printf("{cs} size %zu\\n", sizeof({cs}));')
        for f in fields:
            lines.append(f'  printf("{cs} {f} %zu\\n", offsetof({cs}, {f}));')
    lines += ["  return 0;", "}"]
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "abi.c"), os.path.join(td, "abi")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = {}
    for ln in out.splitlines():
        cs, what, val = ln.split()
        got[(cs, what)] = int(val)
    for cs, pyname, fields in PAIRS:
        cls = getattr(capi, pyname)
        assert ctypes.sizeof(cls) == got[(cs, "size")], (cs, ctypes.sizeof(cls), got[(cs, "size")])
        for f in fields:
            assert getattr(cls, f).offset == got[(cs, f)], (cs, f, getattr(cls, f).offset, got[(cs, f)])


def test_gradient_buffer_bucket_order():
    import torch
    from team_b200 import capi, head
    names = [n for n, _ in head.GRAD_LAYOUT]
    assert names[:4] == ["w_fc", "w_q", "w_k", "w_v"], names          # early buckets of team_head_grads.comm
    assert all(sz % 4 == 0 for _, sz in head.GRAD_LAYOUT)              # float4 granularity of the exchange
    T, ppt = 3, 10
    z = lambda *s: torch.zeros(s)
    pack = head.HeadParamPack([z(512, 512)] * T, [z(512)] * T, [z(512, 512)] * T, [z(512)] * T, [z(512, 512)] * T, [z(512)] * T,
                              [z(ppt, 512)] * T, z(10, 512), z(512, 512), z(512, 512), z(512, 512), z(512, 512), z(512), z(512), z(512))
    n = head.HeadStepRunner.grad_numel(pack)
    assert n == sum(sz for _, sz in head.GRAD_LAYOUT) + T * ppt * capi.D and n % 4 == 0


def test_dynamic_temperature_matches_the_reference_formula():
    import math
    from team_b200 import ops
    assert ops.dynamic_temperature(0.07) == pytest.approx(0.07)
    for e, m in ((0, 10), (3, 10), (10, 10), (7, 20)):
        want = 0.07 * (0.5 + 0.5 * 0.5 * (1.0 + math.cos(math.pi * e / m)))      # models/proof.py:111-116
        assert ops.dynamic_temperature(0.07, e, m) == pytest.approx(want, rel=1e-12)
