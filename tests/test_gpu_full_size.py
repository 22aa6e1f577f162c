"""Parity at BASELINE.json's FULL sizes (configs 3 and 5), where the CPU oracle cannot run: size-independent
properties plus exact spot checks against a chunked float64/float32 torch reference on the device.
(Config 4 at full size is compared with the oracle directly: test_gpu_two_tower.py::test_config4_shape_vs_oracle.)"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _unit(R, D, seed, like=None, mix=0.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(R, D, device=DEV, generator=g)
    if like is not None:
        x = mix * like + x
    return F.normalize(x, dim=1)


def test_infonce_config3_full_size_loss_gradients_and_symmetry():
    """B = 65,536, D = 128 (17 GB similarity matrix, never materialised).  Reference: the same bf16-rounded operands,
    log-sum-exps accumulated in float64 over 4096-row chunks of an fp32 matmul."""
    from ceo_firm_matching.contrastive import info_nce_loss
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        B, D, T = 65_536, 128, 0.07
        f0 = _unit(B, D, 1)
        c0 = _unit(B, D, 2, like=f0, mix=0.7)                   # positives are genuinely similar
        f, c = f0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
        loss = info_nce_loss(f, c, T)
        loss.backward()
        # symmetric in its arguments (contrastive.py:129-138 averages both directions)
        loss_t = info_nce_loss(c0, f0, T)
        assert float(loss_t) == pytest.approx(float(loss), rel=1e-6)
        # chunked reference on the operands the kernel sees
        fq, cq = f0.bfloat16().float(), c0.bfloat16().float()
        CH = 4096
        row_se = torch.zeros(B, dtype=torch.float64, device=DEV)   # sum_j exp((s_ij - 1)/T)
        col_se = torch.zeros(B, dtype=torch.float64, device=DEV)
        for lo in range(0, B, CH):
            e = torch.exp(((fq[lo:lo + CH] @ cq.t()) - 1.0).double() / T)
            row_se[lo:lo + CH] = e.sum(1)
            col_se += e.sum(0)
        diag = (fq.double() * cq.double()).sum(1)
        lse_r, lse_c = torch.log(row_se) + 1.0 / T, torch.log(col_se) + 1.0 / T
        loss_ref = 0.5 * ((lse_r - diag / T).mean() + (lse_c - diag / T).mean())
        assert float(loss) == pytest.approx(float(loss_ref), rel=2e-5)
        # exact gradient rows for a sample of pairs: dF_i = sum_j (p_ij + q_ij) c_j / (2BT) - c_i / (BT)
        rows = torch.tensor([0, 1, 127, 128, 4095, 4096, 32767, 32768, 65534, 65535, 12345, 54321], device=DEV)
        s = fq[rows].double() @ cq.double().t()
        w = torch.exp((s - 1.0) / T) * (1.0 / row_se[rows, None] + 1.0 / col_se[None, :])
        d_ref = (w @ cq.double()) / (2 * B * T) - cq[rows].double() / (B * T)
        got = f.grad[rows].double()
        scale = float(d_ref.abs().max())
        assert float((got - d_ref).abs().max()) <= 6e-3 * scale          # G re-staged as bf16 (see test_gpu_infonce.py)
        s2 = cq[rows].double() @ fq.double().t()
        w2 = torch.exp((s2 - 1.0) / T) * (1.0 / col_se[rows, None] + 1.0 / row_se[None, :])
        d2_ref = (w2 @ fq.double()) / (2 * B * T) - fq[rows].double() / (B * T)
        assert float((c.grad[rows].double() - d2_ref).abs().max()) <= 6e-3 * float(d2_ref.abs().max())
        assert bool(torch.isfinite(f.grad).all()) and bool(torch.isfinite(c.grad).all())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_allpairs_config5_full_size_properties_and_exact_spot_rows():
    """1M CEOs x 1M firms, D = 60, top-100 per CEO: every list sorted / unique / in range / scored exactly, and for a
    sample of rows the index list equals the exact float64 ranking bit for bit."""
    from ceo_firm_matching.scoring import score_topk
    N, D, k, scale = 1_000_000, 60, 100, 1 / 0.07
    ceos, firms = _unit(N, D, 11), _unit(N, D, 12)
    s, idx, flags = score_topk(ceos, firms, k, scale, return_flags=True)
    assert s.shape == (N, k) and idx.shape == (N, k)
    assert int(flags.sum()) == 0                                        # the filter's completeness proof held everywhere
    assert bool((idx >= 0).all()) and bool((idx < N).all())
    assert bool((s[:, :-1] >= s[:, 1:]).all())                          # sorted by score, descending
    srt = idx.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                      # no column twice in a list
    rows = torch.tensor([0, 1, 127, 128, 255, 256, 131071, 131072, 500_000, 999_998, 999_999, 424_242], device=DEV)
    exact = ceos[rows].double() @ firms.double().t() * scale            # [12, 1M] float64
    order = torch.argsort(-exact, dim=1, stable=True)[:, :k]
    assert torch.equal(idx[rows], order)
    assert float((s[rows].double() - torch.gather(exact, 1, order)).abs().max()) <= 2e-6 * float(exact.abs().max())
    # scores of ALL returned pairs recomputed exactly for a stride of rows
    sub = torch.arange(0, N, 997, device=DEV)
    re = (ceos[sub].double().unsqueeze(1) * firms[idx[sub]].double()).sum(-1) * scale
    assert float((s[sub].double() - re).abs().max()) <= 2e-6 * float(re.abs().max())
