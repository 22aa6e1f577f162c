"""Embedding-gradient segment reduce under skewed and low-cardinality indices (data.py:120-135: the reference's real
tables have 2-4 classes; SURVEY 8d asks for Zipf(1.05) indices over 1M rows): runs of tens of thousands of equal keys
must give the dense gradient torch's index_add gives (fp32 summation-order tolerance), bit-identically from run to run,
through the single-tower, the joint and the full model path."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _zipf(n, rows, s, gen):
    """Zipf(s) over [0, rows) by inverse-CDF sampling (rank r has weight (r + 1)^-s)."""
    w = torch.arange(1, rows + 1, dtype=torch.float64).pow(-s)
    cdf = torch.cumsum(w, 0)
    u = torch.rand(n, generator=gen, dtype=torch.float64) * cdf[-1]
    return torch.searchsorted(cdf, u).clamp_(max=rows - 1)


def _reference(x_cat, dx, rows, E):
    outs = []
    for t, n in enumerate(rows):
        g = torch.zeros(n, E, dtype=torch.float64, device=dx.device)
        g.index_add_(0, x_cat[:, t], dx[:, t * E:(t + 1) * E].double())
        outs.append(g)
    return outs


@pytest.mark.parametrize("name,B,rows,E", [
    ("reference cardinalities (firm)", 65536, [4, 4, 2, 2], 48),
    ("reference cardinalities (ceo)", 65536, [2, 4, 2, 2, 2, 2, 2], 8),
    ("one class", 5000, [1, 3], 8),
    ("run lengths around the chunk size", 1000, [15, 16, 17], 12),
    ("zipf(1.05) over 1M rows", 65536, [1000000, 1000000], 48),
    ("scalar slices", 3000, [3, 700], 6),
])
def test_segment_reduce_matches_index_add(name, B, rows, E):
    from ceo_firm_matching import ops
    gen = torch.Generator().manual_seed(11)
    if "zipf" in name:
        x_cat = torch.stack([_zipf(B, n, 1.05, gen) for n in rows], 1)
    else:
        x_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in rows], 1)
    dx = torch.randn(B, len(rows) * E, generator=gen)
    x_cat, dx = x_cat.to(DEV), dx.to(DEV)
    embs = torch.nn.ModuleList([torch.nn.Embedding(n, E) for n in rows]).to(DEV)
    h = ops.TowerHandle(embs, torch.nn.Sequential(), (), (), (), 0)
    got = ops.embedding_grads(h, x_cat, dx, B)
    got2 = ops.embedding_grads(h, x_cat, dx, B)
    want = _reference(x_cat, dx, rows, E)
    for t in range(len(rows)):
        assert torch.equal(got[t], got2[t]), f"{name}: table {t} differs between two runs"
        scale = want[t].abs().max().item()
        # fp32 sums of up to B terms in a fixed order: error <= ~sqrt(run length) * 2^-24 * scale of the terms
        err = (got[t].double() - want[t]).abs().max().item()
        assert err <= 3e-6 * scale + 1e-6, f"{name}: table {t} off by {err:.3e} (scale {scale:.3e})"
        untouched = torch.ones(rows[t], dtype=torch.bool, device=DEV)
        untouched[x_cat[:, t]] = False
        assert not got[t][untouched].any(), f"{name}: rows nobody indexed are not zero"


@pytest.mark.parametrize("cards", [([4, 4, 2, 2], [2, 4, 2, 2, 2, 2, 2])])
def test_model_step_with_reference_cardinalities_at_full_batch(cards):
    """Whole train step at B = 65 536 with the reference's real table sizes: the joint sort + chunked reduce path."""
    import oracle
    from ceo_firm_matching import CEOFirmMatcher, Config
    from helpers import check_grads, load_into
    f_cards, c_cards = cards
    B = 65536
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=5)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    gen = torch.Generator().manual_seed(6)
    ins = [torch.randn(B, 12, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1),
           torch.randn(B, 2, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1),
           torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5]
    grads = []
    for _ in range(2):
        m = load_into(CEOFirmMatcher(meta, Config()), p).to(DEV).train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.use_persistent_table_grads(True)
        loss, _ = m.forward_loss(*[x.to(DEV) for x in ins])
        loss.backward()
        grads.append({k: q.grad.clone() for k, q in m.named_parameters()})
    for k in grads[0]:
        assert torch.equal(grads[0][k], grads[1][k]), f"{k} differs between two runs"
    po = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
    lo = oracle.weighted_mse(oracle.two_tower_forward(po, *ins[:4], training=True), ins[4], ins[5])
    lo.backward()
    assert abs(float(loss) - float(lo)) <= 2e-5 * abs(float(lo))
    check_grads(m, {k: po[k].grad for k, _ in m.named_parameters()}, 2e-4, "low-cardinality step")
