"""Host logic of the frozen-tower feature cache (team_b200/feature_cache.py) - CPU: a counting fake tower
stands for CLIP.  Checked: one image-tower pass for the three calls the learner makes per batch (models/proof.py:418,
:425, :430), the text tower only ever sees unseen rows, results identical to the uncached tower, invalidation when the tower
changes, bypass while the tower is trainable."""
import torch
import torch.nn as nn

from team_b200 import feature_cache as fc


class CountingTower(nn.Module):
    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.randn(16, 512), requires_grad=False)
        self.emb = nn.Parameter(torch.randn(50, 512), requires_grad=False)
        self.logit_scale = nn.Parameter(torch.tensor(2.0))
        self.image_rows = self.text_rows = 0

    def encode_image(self, x, normalize=False):
        self.image_rows += x.shape[0]
        return x @ self.w

    def encode_text(self, tok, normalize=False):
        self.text_rows += tok.shape[0]
        return self.emb[tok].sum(1)


def test_three_image_calls_one_tower_pass():
    t = CountingTower(); c = fc.TowerCache(t)
    x = torch.randn(8, 16)
    a, b, d = c.image(x), c.image(x), c.image(x)
    assert t.image_rows == 8 and torch.equal(a, x @ t.w) and a is b is d
    y = torch.randn(8, 16)
    assert torch.equal(c.image(y), y @ t.w) and t.image_rows == 16
    x.add_(1.0)                                    # in-place change of a cached input: version counter differs -> recompute
    assert torch.equal(c.image(y), y @ t.w) and t.image_rows == 16
    assert torch.equal(c.image(x), x @ t.w) and t.image_rows == 24


def test_text_tower_sees_each_row_once():
    t = CountingTower(); c = fc.TowerCache(t)
    cls_tok = torch.randint(0, 50, (6, 7))
    full = c.text(cls_tok)
    assert torch.equal(full, t.encode_text(cls_tok)); t.text_rows = 0
    for step in range(5):
        y = torch.randint(0, 6, (32,))
        out = c.text(cls_tok[y])                   # per-sample prompts of a batch: <= 6 distinct rows, all seen before
        assert torch.equal(out, full[y])
    assert t.text_rows == 0
    new = torch.randint(0, 50, (3, 7))
    mixed = torch.cat([cls_tok[:2], new])
    ref = t.emb[mixed].sum(1)
    t.text_rows = 0
    assert torch.equal(c.text(mixed), ref) and t.text_rows <= 3
    assert c.stats["text_tower_rows"] <= 6 + 3


def test_cache_follows_the_tower():
    t = CountingTower(); c = fc.TowerCache(t)
    x, tok = torch.randn(4, 16), torch.randint(0, 50, (4, 7))
    c.image(x); c.text(tok)
    with torch.no_grad():
        t.w.mul_(2.0); t.emb.add_(1.0)             # e.g. load_state_dict: parameter versions change
    assert torch.equal(c.image(x), x @ t.w) and torch.equal(c.text(tok), t.emb[tok].sum(1))
    t.w.requires_grad_(True)                       # trainable tower: no caching, gradients flow
    n0 = t.image_rows
    out = c.image(x); c.image(x)
    assert out.requires_grad and t.image_rows == n0 + 8


def test_indexed_store():
    t = CountingTower()
    st = fc.IndexedFeatureStore(lambda v: t.encode_image(v), capacity=100, device="cpu")
    data = torch.randn(100, 16)
    for epoch in range(3):
        perm = torch.randperm(100, generator=torch.Generator().manual_seed(epoch))
        for i in range(0, 100, 25):
            idx = perm[i:i + 25]
            assert torch.allclose(st.get(idx, data[idx]), data[idx] @ t.w)
    assert t.image_rows == 100 and st.tower_rows == 100          # one tower pass per sample for the whole task
    st.clear()
    st.get(torch.arange(10), data[:10])
    assert t.image_rows == 110


def test_row_match_in_slabs_equals_one_shot():
    from team_b200.feature_cache import _match_rows
    g = torch.Generator().manual_seed(3)
    table = torch.randint(0, 50, (37, 7), generator=g)
    rows = torch.cat([table[torch.randperm(37, generator=g)[:20]], torch.randint(50, 60, (9, 7), generator=g)])
    want = torch.tensor([next((j for j in range(37) if torch.equal(table[j], r)), -1) for r in rows])
    assert torch.equal(_match_rows(rows, table), want)
    assert torch.equal(_match_rows(rows, table, max_cells=37 * 7 * 3), want)        # 3 rows per slab
    assert torch.equal(_match_rows(rows, table, max_cells=1), want)                  # 1 row per slab
