"""GPU parity of all-pairs scoring + top-k: index lists must equal the exact ranking of the fp32 operands
(oracle.allpairs_topk forms scores in float64), bit for bit, including order and tie-breaks."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from helpers import assert_close_scaled, load_golden, t

pytestmark = pytest.mark.gpu
DEV = "cuda"
SCALE = float(np.exp(np.log(1 / 0.07)))


def _unit(R, D, seed):
    g = torch.Generator().manual_seed(seed)
    return F.normalize(torch.randn(R, D, generator=g), dim=1)


def test_matches_reference_golden_ranking():
    from ceo_firm_matching.scoring import score_topk
    g = load_golden("allpairs_50x70_d60")
    u, v, scale = t(g["u"]).to(DEV), t(g["v"]).to(DEV), float(g["scale"])
    s, idx = score_topk(u, v, 10, scale)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["ranking"][:, :10])         # np.argsort(-scores)[:10]
    assert_close_scaled(s, np.take_along_axis(g["scores"], g["ranking"][:, :10], 1), 2e-6, "scores")
    s_all, idx_all = score_topk(u, v, 70, scale)                                     # the full ranking (k == C)
    np.testing.assert_array_equal(idx_all.cpu().numpy(), g["ranking"])


@pytest.mark.parametrize("R,C,D,k", [(200, 1000, 60, 10), (300, 5000, 60, 100), (129, 12345, 60, 100),
                                     (1000, 20000, 128, 100), (64, 40000, 30, 128), (2000, 130, 60, 100)])
def test_index_lists_equal_exact_ranking(R, C, D, k):
    from ceo_firm_matching.scoring import score_topk
    u, v = _unit(R, D, 1), _unit(C, D, 2)
    s, idx, flags = score_topk(u.to(DEV), v.to(DEV), k, SCALE, return_flags=True)
    so, io = oracle.allpairs_topk(u, v, k, SCALE)
    np.testing.assert_array_equal(idx.cpu().numpy(), io.numpy())
    assert_close_scaled(s, so, 2e-6, "scores")
    assert int(flags.sum()) == 0                      # random unit vectors: the filter proves completeness everywhere


def test_ties_duplicates_and_short_rows():
    from ceo_firm_matching.scoring import score_topk
    u = _unit(70, 60, 3)
    base = _unit(500, 60, 4)
    v = torch.cat([base, base[:300], base[:100]])      # exact duplicate columns -> exact score ties
    s, idx = score_topk(u.to(DEV), v.to(DEV), 50, 2.0)
    so, io = oracle.allpairs_topk(u, v, 50, 2.0)
    np.testing.assert_array_equal(idx.cpu().numpy(), io.numpy())      # smaller index first among equal scores
    s2, idx2 = score_topk(u.to(DEV), v[:7].to(DEV), 20, 1.0)           # fewer columns than k
    assert idx2.shape == (70, 20) and bool((idx2[:, 7:] == -1).all()) and bool(torch.isinf(s2[:, 7:]).all())
    so2, io2 = oracle.allpairs_topk(u, v[:7], 20, 1.0)
    np.testing.assert_array_equal(idx2[:, :7].cpu().numpy(), io2.numpy())


def test_clustered_columns_force_exact_fallback():
    """Thousands of near-identical columns defeat the filter's proof; flagged rows are redone exactly."""
    from ceo_firm_matching.scoring import score_topk
    g = torch.Generator().manual_seed(9)
    u = _unit(40, 60, 5)
    centre = _unit(1, 60, 6)
    v = F.normalize(centre + 1e-4 * torch.randn(3000, 60, generator=g), dim=1)
    v = torch.cat([v, _unit(2000, 60, 7)])
    s, idx, flags = score_topk(u.to(DEV), v.to(DEV), 100, 1.0, return_flags=True)
    so, io = oracle.allpairs_topk(u, v, 100, 1.0)
    np.testing.assert_array_equal(idx.cpu().numpy(), io.numpy())
    assert int(flags.sum()) > 0


def test_sharded_columns_merge_equals_full():
    from ceo_firm_matching.scoring import merge_topk, score_topk
    u, v = _unit(500, 60, 11).to(DEV), _unit(9000, 60, 12).to(DEV)
    full_s, full_i = score_topk(u, v, 100, SCALE)
    parts = [score_topk(u, v[a:b], 100, SCALE, col_offset=a, return_f64=True)
             for a, b in ((0, 2000), (2000, 2100), (2100, 9000))]
    ms, mi = merge_topk(torch.stack([p[2] for p in parts]), torch.stack([p[1] for p in parts]))   # fp64 scores
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    ms32, mi32 = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert float((mi32 == full_i).float().mean()) > 0.999          # fp32 merge: exact up to fp32-rounding ties


def test_diagonal_ranks_and_retrieval_metrics():
    from ceo_firm_matching.scoring import diagonal_ranks
    f = _unit(700, 60, 21)
    c = F.normalize(f + 0.7 * _unit(700, 60, 22), dim=1)
    ranks = diagonal_ranks(f.to(DEV), c.to(DEV)).cpu().numpy()
    sim = f.double() @ c.double().t()
    order = torch.argsort(-sim, dim=1, stable=True)
    want = (order == torch.arange(700).unsqueeze(1)).float().argmax(1).numpy() + 1
    np.testing.assert_array_equal(ranks, want)
    assert oracle.retrieval_metrics(ranks)["recall@10"] == oracle.retrieval_metrics(want)["recall@10"]


@pytest.mark.parametrize("rb", [1, 2])
def test_row_block_variants_give_exact_ranking(rb):
    """Both CTA shapes of the filter kernel (one / two 128-row blocks, i.e. split vs single candidate stream per row)."""
    from ceo_firm_matching import _native as N
    from ceo_firm_matching.scoring import score_topk
    u, v = _unit(700, 60, 31), _unit(30000, 60, 32)
    N.check(N.lib().cfm_simtile_set_rb(rb))
    try:
        s, idx, flags = score_topk(u.to(DEV), v.to(DEV), 100, SCALE, return_flags=True)
    finally:
        N.check(N.lib().cfm_simtile_set_rb(0))
    so, io = oracle.allpairs_topk(u, v, 100, SCALE)
    np.testing.assert_array_equal(idx.cpu().numpy(), io.numpy())
    assert int(flags.sum()) == 0


def test_generate_counterfactuals_matches_a_full_sort_restatement():
    """analytical_extensions.py:405-523 routed to score_topk / target_ranks: the table equals what the reference's
    torch.mm + np.argsort loop gives on the same (device-computed) unit latents."""
    import contextlib
    import io
    import pandas as pd
    import oracle
    from ceo_firm_matching import CEOFirmMatcher, Config
    from ceo_firm_matching.analytical_extensions import generate_counterfactuals, _unit_latents
    from helpers import load_into
    f_cards, c_cards = [9, 5, 3, 2], [2, 4, 3, 2, 2, 5, 2]
    n = 600
    gen = torch.Generator().manual_seed(21)
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=8)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    model = load_into(CEOFirmMatcher(meta, Config()), p).to(DEV)
    data = {"firm_numeric": torch.randn(n, 12, generator=gen),
            "firm_cat": torch.stack([torch.randint(0, c, (n,), generator=gen) for c in f_cards], 1),
            "ceo_numeric": torch.randn(n, 2, generator=gen),
            "ceo_cat": torch.stack([torch.randint(0, c, (n,), generator=gen) for c in c_cards], 1)}
    rng = np.random.default_rng(3)
    df = pd.DataFrame({"gvkey": rng.integers(0, 150, n), "match_exec_id": rng.integers(0, 400, n),
                       "fiscalyear": rng.integers(2000, 2020, n), "match_means": rng.normal(size=n)})
    with contextlib.redirect_stdout(io.StringIO()):
        cf = generate_counterfactuals(model, data, df, device=DEV)
    # restatement of the reference's loop on the same latents
    with torch.no_grad():
        model.eval()
        u, v = _unit_latents(model, data, torch.device(DEV))
    u, v = u.cpu().double(), v.cpu().double()
    scale = float(model.logit_scale.exp())
    d = df.reset_index(drop=True)
    firm_idx = d.loc[d.groupby("gvkey")["fiscalyear"].idxmax()].index.values[:200]
    ceo_idx = d.drop_duplicates(subset="match_exec_id", keep="last").index.values[:1000]
    scores = (u[firm_idx] @ v[ceo_idx].t()).numpy() * scale
    assert len(cf) == len(firm_idx)
    for i, fi in enumerate(firm_idx):
        order = np.argsort(-scores[i], kind="stable")
        row = cf.iloc[i]
        assert row["firm_id"] == d.loc[fi, "gvkey"]
        assert row["best_ceo"] == d.loc[ceo_idx[order[0]], "match_exec_id"]
        assert row["worst_ceo"] == d.loc[ceo_idx[order[-1]], "match_exec_id"]
        assert row["best_score"] == pytest.approx(scores[i, order[0]], rel=2e-6, abs=1e-6)
        assert row["worst_score"] == pytest.approx(scores[i, order[-1]], rel=2e-6, abs=1e-6)
        hit = (d.loc[ceo_idx, "match_exec_id"] == d.loc[fi, "match_exec_id"]).values
        if hit.any():
            assert row["actual_rank"] == int((order == hit.argmax()).argmax()) + 1
        else:
            assert row["actual_rank"] is None or np.isnan(row["actual_rank"])
