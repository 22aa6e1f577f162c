"""Measured error budget of the fp32-class two-tower path ("tighten the fp32 tolerance or publish measured
per-tensor errors").  Ground truth is the oracle run in float64; next to the CUDA path (3xTF32 tensor-core products,
fp32 everywhere else) the same oracle in float32 — i.e. stock ATen fp32 — is measured against that truth, so the table
says how far two honest fp32 evaluations of this network are from each other.

Two figures per tensor, both relative:  max|a - e| / max|e|  (error against the tensor's scale, what the parity tests
bound) and the relative L2 error  ||a - e|| / ||e||.  The test asserts the north-star's fp32 bar, 1e-5 relative to the
tensor's scale, for the scores, the loss and EVERY gradient (measured on B200: <= 2.7e-6, torch fp32 itself <= 1.4e-6),
and that the CUDA path stays within 4x of torch fp32's own error (plus a floor of 2e-6); the table is written to
``gpurun_out/r02_error_budget.md`` (copied to ``profiles/``)."""
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import ROOT, load_into

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(params, f_cards, c_cards):
    from ceo_firm_matching import CEOFirmMatcher, Config
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    m = CEOFirmMatcher(meta, Config())
    load_into(m, params)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m.to(DEV).train()


def _oracle_run(p, ins, dtype):
    cast = lambda v: v.to(dtype) if v.is_floating_point() else v
    po = {k: cast(v.clone()).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
    f_num, f_cat, c_num, c_cat, target, weights = [cast(x) for x in ins]
    preds = oracle.two_tower_forward(po, f_num, f_cat, c_num, c_cat, training=True)
    loss = oracle.weighted_mse(preds, target, weights)
    loss.backward()
    grads = {k: v.grad.double() for k, v in po.items() if v.requires_grad and v.grad is not None}
    return preds.detach().double(), loss.detach().double(), grads


def _errs(a, e):
    a, e = np.asarray(a, dtype=np.float64), np.asarray(e, dtype=np.float64)
    scale = max(float(np.abs(e).max()), 1e-300)
    return float(np.abs(a - e).max() / scale), float(np.linalg.norm(a - e) / max(np.linalg.norm(e), 1e-300))


@pytest.mark.parametrize("B", [4096])
def test_error_budget_vs_float64(B):
    f_cards, c_cards = [50, 20, 9, 4], [2, 4, 30, 2, 2, 5, 2]
    gen = torch.Generator().manual_seed(2024)
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=7)
    f_num, c_num = torch.randn(B, 12, generator=gen), torch.randn(B, 2, generator=gen)
    f_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1)
    c_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1)
    target = torch.randn(B, 1, generator=gen)
    weights = 1.0 / (torch.rand(B, 1, generator=gen) * 0.9 + 0.1) ** 2
    ins = (f_num, f_cat, c_num, c_cat, target, weights)

    preds64, loss64, g64 = _oracle_run(p, ins, torch.float64)
    preds32, loss32, g32 = _oracle_run(p, ins, torch.float32)

    m = _model(p, f_cards, c_cards)
    loss, preds = m.forward_loss(*[x.to(DEV) for x in ins])
    loss.backward()
    gcu = {k: v.grad.detach().cpu().double() for k, v in m.named_parameters() if v.grad is not None}

    dead = {k for k in g64 if float(g64[k].abs().max()) < 1e-12 * max(float(v.abs().max()) for v in g64.values())}
    rows = [("scores", _errs(preds.detach().cpu().double(), preds64), _errs(preds32, preds64)),
            ("loss", _errs(loss.detach().cpu().double(), loss64), _errs(loss32, loss64))]
    for k in sorted(g64):
        if k in dead or k not in gcu:
            continue                                  # mathematically-zero gradients (Linear bias before train-mode BN)
        rows.append((f"grad {k}", _errs(gcu[k], g64[k]), _errs(g32[k], g64[k])))

    lines = ["| tensor | CUDA 3xTF32: max err / scale | rel. L2 | torch fp32 (oracle): max err / scale | rel. L2 |",
             "|---|---:|---:|---:|---:|"]
    for name, (a_max, a_l2), (t_max, t_l2) in rows:
        lines.append(f"| `{name}` | {a_max:.2e} | {a_l2:.2e} | {t_max:.2e} | {t_l2:.2e} |")
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "r02_error_budget.md"), "w") as f:
        f.write(f"Two-tower fwd + weighted MSE + bwd, B = {B}, dropout off, train-mode BatchNorm; truth = oracle in float64.\n\n")
        f.write("\n".join(lines) + "\n")

    for name, (a_max, a_l2), (t_max, t_l2) in rows:
        bar = 1e-5
        assert a_max <= bar, f"{name}: max error / scale {a_max:.2e} above the fp32 bar {bar:.0e}"
        assert a_max <= 4 * t_max + 2e-6, f"{name}: CUDA {a_max:.2e} vs torch fp32's own {t_max:.2e}"
