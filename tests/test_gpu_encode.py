"""§8(f4): the reusable encode() op on a module with the reference's extension-model layout
(multitask_model.py:60-147: towers + task heads on the concatenated embeddings) and batched grid scoring
(visualization.py:99-123).  The expected values come from the module's own torch path — the reference's encode body
executed by stock torch ops (fp32, TF32 matmuls off) on the same device."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import assert_close_scaled

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tower(i, o):
    return nn.Sequential(nn.Linear(i, 64), nn.BatchNorm1d(64), nn.ReLU(), nn.Dropout(0.0),
                         nn.Linear(64, 32), nn.BatchNorm1d(32), nn.ReLU(), nn.Dropout(0.0), nn.Linear(32, o))


class MultiTaskLike(nn.Module):
    """Attribute layout of the reference's MultiTaskCEOFirmMatcher (multitask_model.py:60-125)."""

    def __init__(self, f_cards, c_cards, latent=60):
        super().__init__()
        self.firm_embeddings = nn.ModuleList(nn.Embedding(n, 48) for n in f_cards)
        self.ceo_embeddings = nn.ModuleList(nn.Embedding(n, 8) for n in c_cards)
        self.firm_tower = _tower(12 + 48 * len(f_cards), latent)
        self.ceo_tower = _tower(2 + 8 * len(c_cards), latent)
        self.comp_head = nn.Sequential(nn.Linear(2 * latent, 32), nn.ReLU(), nn.Linear(32, 1))

    def encode_torch(self, f_num, f_cat, c_num, c_cat):          # multitask_model.py:127-147 verbatim in structure
        f = torch.cat([f_num] + [e(f_cat[:, i]) for i, e in enumerate(self.firm_embeddings)], 1)
        c = torch.cat([c_num] + [e(c_cat[:, i]) for i, e in enumerate(self.ceo_embeddings)], 1)
        return F.normalize(self.firm_tower(f), dim=1), F.normalize(self.ceo_tower(c), dim=1)


def _inputs(B, f_cards, c_cards, seed):
    g = torch.Generator().manual_seed(seed)
    f_num, c_num = torch.randn(B, 12, generator=g), torch.randn(B, 2, generator=g)
    f_cat = torch.stack([torch.randint(0, n, (B,), generator=g) for n in f_cards], 1)
    c_cat = torch.stack([torch.randint(0, n, (B,), generator=g) for n in c_cards], 1)
    return [x.to(DEV) for x in (f_num, f_cat, c_num, c_cat)]


@pytest.mark.parametrize("training", [False, True])
def test_encode_matches_the_modules_own_torch_path(training):
    from ceo_firm_matching.encode import encode
    torch.backends.cuda.matmul.allow_tf32 = False
    f_cards, c_cards = [9, 5, 3, 2], [2, 4, 3, 2, 2, 5, 2]
    torch.manual_seed(3)
    m = MultiTaskLike(f_cards, c_cards).to(DEV).train(training)
    ins = _inputs(300, f_cards, c_cards, 11)
    target = torch.randn(300, 1, device=DEV)

    def head_loss(u, v):
        return ((m.comp_head(torch.cat([u, v], 1)) - target) ** 2).mean()

    sd = {k: v.clone() for k, v in m.state_dict().items()}
    u0, v0 = m.encode_torch(*ins)
    l0 = head_loss(u0, v0); l0.backward()
    g0 = {k: p.grad.clone() for k, p in m.named_parameters()}
    stats0 = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
    m.load_state_dict(sd); m.zero_grad(set_to_none=True)

    u1, v1 = encode(m, *ins)
    l1 = head_loss(u1, v1); l1.backward()
    assert_close_scaled(u1, u0, 2e-5, "firm embedding"); assert_close_scaled(v1, v0, 2e-5, "ceo embedding")
    assert_close_scaled(l1, l0, 2e-5, "head loss")
    gmax = max(float(g.abs().max()) for g in g0.values())
    for k, p in m.named_parameters():
        if training and k.endswith(("tower.0.bias", "tower.4.bias")):
            continue                         # bias before a train-mode BatchNorm: true gradient 0, noise on both sides
        assert_close_scaled(p.grad, g0[k], 5e-5, f"grad {k}", floor=2e-6 * gmax)
    if training:                             # BatchNorm running statistics advance exactly once
        for k, v in m.state_dict().items():
            if "running" in k:
                assert_close_scaled(v.float(), stats0[k].float(), 2e-5, k)


def test_encode_rejects_other_tower_layouts():
    from ceo_firm_matching.encode import encode
    m = MultiTaskLike([3], [2]).to(DEV)
    m.firm_tower = nn.Sequential(nn.Linear(60, 64), nn.ReLU(), nn.Linear(64, 60)).to(DEV)
    with pytest.raises(ValueError):
        encode(m, *_inputs(4, [3], [2], 0))


def test_score_grid_equals_the_cell_by_cell_loop():
    from ceo_firm_matching import CEOFirmMatcher, Config
    from ceo_firm_matching.encode import score_grid
    f_cards, c_cards = [9, 5, 3, 2], [2, 4, 3, 2, 2, 5, 2]
    torch.manual_seed(5)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    m = CEOFirmMatcher(meta, Config()).to(DEV).eval()
    base = _inputs(1, f_cards, c_cards, 2)
    x_vals, y_vals = np.linspace(-2, 2, 7), [0, 1, 2, 3]
    heat = score_grid(m, base, ("firm_numeric", 3), x_vals, ("ceo_cat", 5), y_vals)
    assert heat.shape == (4, 7)
    with torch.no_grad():                    # visualization.py:99-123: one single-row forward per cell
        for i, yv in enumerate(y_vals):
            for j, xv in enumerate(x_vals):
                f_num, f_cat, c_num, c_cat = [t.clone() for t in base]
                f_num[:, 3] = float(xv); c_cat[:, 5] = int(yv)
                assert float(heat[i, j]) == pytest.approx(float(m(f_num, f_cat, c_num, c_cat)), rel=2e-5, abs=2e-5)
