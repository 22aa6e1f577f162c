"""Rank of the positive counted in the epilogue of the tcgen05 similarity kernel (``cfm_allpairs_diag_rank`` behind
``scoring.diagonal_ranks(method="tensor")``; contrastive.py:296-332, run_deep_extensions.py:470-483): ranks must be
EQUAL to the exact fp64 SIMT kernel's and to a stable descending sort of the float64 score matrix, ties included."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _unit(R, D, seed):
    g = torch.Generator().manual_seed(seed)
    return F.normalize(torch.randn(R, D, generator=g), dim=1)


def _sorted_ranks(u, v, offset=0):
    """1 + position of column i + offset in the stable descending sort of row i (float64 scores); C if absent."""
    sim = u.double() @ v.double().t()
    order = torch.argsort(-sim, dim=1, stable=True)
    tgt = torch.arange(u.shape[0]).unsqueeze(1) + offset
    hit = order == tgt
    pos = hit.float().argmax(1) + 1
    return torch.where(hit.any(1), pos, torch.full_like(pos, v.shape[0])).numpy()


@pytest.mark.parametrize("rb", [1, 2])
@pytest.mark.parametrize("shape", [(3000, 3000, 60), (2500, 4100, 60), (1500, 1300, 60), (2100, 2100, 128), (2304, 2304, 30)])
def test_tensor_ranks_equal_exact_and_sorted(shape, rb):
    from ceo_firm_matching import _native as N
    from ceo_firm_matching.scoring import diagonal_ranks
    R, C, D = shape
    u = _unit(R, D, R + D)
    v = _unit(C, D, C + 7)
    n = min(R, C)
    v[:n] = F.normalize(0.6 * u[:n] + v[:n], dim=1)       # positives well above the bulk, but not always rank 1
    N.check(N.lib().cfm_simtile_set_rb(rb))
    try:
        got = diagonal_ranks(u.to(DEV), v.to(DEV), method="tensor").cpu().numpy()
    finally:
        N.check(N.lib().cfm_simtile_set_rb(0))
    exact = diagonal_ranks(u.to(DEV), v.to(DEV), method="exact").cpu().numpy()
    np.testing.assert_array_equal(got, exact)
    np.testing.assert_array_equal(got, _sorted_ranks(u, v))
    assert 1 < np.median(got) or (got == 1).mean() < 1.0   # the case is not trivially all-ones


def test_random_pairs_many_near_ties_and_exact_duplicates():
    """Untrained-model regime (positives inside the bulk: thousands of listed pairs) plus exact duplicates of the
    positive column on both sides of it (ties go to the smaller column index)."""
    from ceo_firm_matching.scoring import diagonal_ranks
    R = 2600
    u, v = _unit(R, 60, 1), _unit(R, 60, 2)
    v[10] = v[2000]; v[2500] = v[2000]                    # row 2000's positive has a twin before and after it
    v[5] = v[7]                                           # rows 5 and 7: each other's positive is a twin
    got = diagonal_ranks(u.to(DEV), v.to(DEV), method="tensor").cpu().numpy()
    np.testing.assert_array_equal(got, _sorted_ranks(u, v))
    np.testing.assert_array_equal(got, diagonal_ranks(u.to(DEV), v.to(DEV), method="exact").cpu().numpy())


def test_row_blocks_overflow_fallback_and_bf16_operands(monkeypatch):
    from ceo_firm_matching import scoring
    R = 2300
    u, v = _unit(R, 60, 3), _unit(R, 60, 4)
    v = F.normalize(0.3 * u + v, dim=1)
    want = _sorted_ranks(u, v)
    monkeypatch.setattr(scoring, "_RANK_BLOCK_ROWS", 1024)            # three row blocks with their own diagonal offsets
    np.testing.assert_array_equal(scoring.diagonal_ranks(u.to(DEV), v.to(DEV), method="tensor").cpu().numpy(), want)
    monkeypatch.setattr(scoring, "_RANK_AMB_CAP", 16)                 # pair list overflows -> exact kernel for the block
    np.testing.assert_array_equal(scoring.diagonal_ranks(u.to(DEV), v.to(DEV), method="tensor").cpu().numpy(), want)
    monkeypatch.undo()
    # untrained embeddings (positives inside the bulk): a list too small for a 1024-row block, large enough for its
    # quarters -> the blocks are split and run again on the tensor path, the exact kernel is never called
    ur, vr = _unit(R, 60, 13), _unit(R, 60, 14)
    monkeypatch.setattr(scoring, "_RANK_BLOCK_ROWS", 1024)
    monkeypatch.setattr(scoring, "_RANK_AMB_CAP", 6000)
    monkeypatch.setattr(scoring, "_RANK_MIN_BLOCK", 128)

    def no_exact(*a, **k):
        raise AssertionError("exact kernel called")

    monkeypatch.setattr(scoring, "target_ranks", no_exact)
    np.testing.assert_array_equal(scoring.diagonal_ranks(ur.to(DEV), vr.to(DEV), method="tensor").cpu().numpy(),
                                  _sorted_ranks(ur, vr))
    monkeypatch.undo()
    big = 4000.0                                                      # outside fp16's comfortable range -> bf16 copies
    np.testing.assert_array_equal(scoring.diagonal_ranks((big * u).to(DEV), v.to(DEV), method="tensor").cpu().numpy(),
                                  _sorted_ranks(big * u, v))


def test_auto_switches_to_the_tensor_path_and_metrics_agree():
    """compute_retrieval_metrics' window (5 000 rows = 25M pairs) takes the tensor path under method="auto"."""
    import oracle
    from ceo_firm_matching import _native as N
    from ceo_firm_matching.scoring import diagonal_ranks
    u = _unit(5000, 60, 5)
    v = F.normalize(0.8 * u + _unit(5000, 60, 6), dim=1)
    before = N.lib().cfm_launch_count(0)
    got = diagonal_ranks(u.to(DEV), v.to(DEV)).cpu().numpy()
    assert N.lib().cfm_launch_count(0) - before >= 6                   # pack x2, prepare, simtile, resolve, finish
    want = _sorted_ranks(u, v)
    np.testing.assert_array_equal(got, want)
    assert oracle.retrieval_metrics(got) == oracle.retrieval_metrics(want)


def test_degenerate_inputs_every_pair_a_tie():
    """All rows and all columns identical: every score ties with the positive, every pair falls inside the window (the
    64-bit pair counter overflows the list by orders of magnitude), ranks are decided by the column index alone."""
    from ceo_firm_matching.scoring import diagonal_ranks
    R = 4300
    one = F.normalize(torch.randn(1, 60, generator=torch.Generator().manual_seed(9)), dim=1)
    u, v = one.repeat(R, 1), one.repeat(R, 1)
    got = diagonal_ranks(u.to(DEV), v.to(DEV), method="tensor").cpu().numpy()
    np.testing.assert_array_equal(got, np.arange(1, R + 1))
