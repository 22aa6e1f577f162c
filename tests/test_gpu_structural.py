"""GPU parity of the structural distillation path (towers -> softmax.A.softmax + KL) against the reference's
golden vectors, the reference's own known-answer test, and the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from helpers import assert_close_scaled, check_grads, load_golden, load_into, params_from, t

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5
META = {"n_firm_num": 12, "n_ceo_num": 2, "firm_cat_cards": [4, 4, 2, 2], "ceo_cat_cards": [2, 4, 2, 2, 2, 2, 2]}


def _model(params=None, meta=META):
    from ceo_firm_matching import StructuralConfig, StructuralDistillationNet
    m = StructuralDistillationNet(meta, StructuralConfig())
    if params is not None:
        load_into(m, params)
    return m.to(DEV)


def _zero_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0


@pytest.mark.parametrize("name", ["structural_b29", "structural_b200"])
def test_matches_reference_golden(name):
    g = load_golden(name)
    p = params_from(g)
    m = _model(p)
    assert set(m.state_dict().keys()) == set(p.keys())
    f_num, f_cat, c_num, c_cat = [t(g[k]).to(DEV) for k in ("f_num", "f_cat", "c_num", "c_cat")]
    # eval forward + input sensitivities (IlluminationEngine, structural_explain.py:68-89)
    m.eval()
    fn, cn = f_num.clone().requires_grad_(True), c_num.clone().requires_grad_(True)
    c_logits, f_logits, match = m(fn, f_cat, cn, c_cat)
    match.sum().backward()
    assert_close_scaled(c_logits, g["eval_c_logits"], TOL, "eval c_logits")
    assert_close_scaled(f_logits, g["eval_f_logits"], TOL, "eval f_logits")
    assert_close_scaled(match, g["eval_match"], TOL, "eval match")
    assert_close_scaled(fn.grad, g["eval_dmatch_df_num"], 5e-5, "dmatch/df_num")
    assert_close_scaled(cn.grad, g["eval_dmatch_dc_num"], 5e-5, "dmatch/dc_num")
    m.zero_grad()
    # train step with the reference's own loss expression (structural_training.py:75-77)
    _zero_dropout(m)
    m.train()
    tc, tf = t(g["target_ceo"]).to(DEV), t(g["target_firm"]).to(DEV)
    c_logits, f_logits, match = m(f_num, f_cat, c_num, c_cat)
    crit = torch.nn.KLDivLoss(reduction="batchmean")
    loss = crit(F.log_softmax(c_logits, 1), tc) + crit(F.log_softmax(f_logits, 1), tf)
    loss.backward()
    assert_close_scaled(loss, g["train_loss"], TOL, "loss (torch KL on fused logits)")
    assert_close_scaled(match, g["train_match"], TOL, "train match")
    check_grads(m, {k[5:]: v for k, v in g.items() if k.startswith("grad/")}, 5e-5)
    sd = m.state_dict()
    for k, v in g.items():
        if k.startswith("after/"):
            assert_close_scaled(sd[k[6:]].float(), np.asarray(v, dtype=np.float64), TOL, k)


@pytest.mark.parametrize("name", ["structural_b29", "structural_b200"])
def test_fused_kl_loss_matches_golden(name):
    g = load_golden(name)
    m = _model(params_from(g))
    _zero_dropout(m)
    m.train()
    f_num, f_cat, c_num, c_cat, tc, tf = [t(g[k]).to(DEV) for k in
                                          ("f_num", "f_cat", "c_num", "c_cat", "target_ceo", "target_firm")]
    c_logits, f_logits = m.logits(f_num, f_cat, c_num, c_cat)
    loss = m.distillation_loss(c_logits, f_logits, tc, tf)       # includes a t == 0 row (xlogy edge case)
    loss.backward()
    assert_close_scaled(loss, g["train_loss"], TOL, "fused KL loss")
    check_grads(m, {k[5:]: v for k, v in g.items() if k.startswith("grad/")}, 5e-5)


def test_reference_known_answer_bilinear():
    """Reference tests/test_structural_model.py:158-178 and :78-101, run on the CUDA build."""
    torch.manual_seed(0)
    m = _model().eval()
    B = 16
    f_num, c_num = torch.randn(B, 12, device=DEV), torch.randn(B, 2, device=DEV)
    f_cat = torch.randint(0, 2, (B, 4), device=DEV)
    c_cat = torch.randint(0, 2, (B, 7), device=DEV)
    with torch.no_grad():
        c_logits, f_logits, match = m(f_num, f_cat, c_num, c_cat)
        pi, q = F.softmax(c_logits, 1), F.softmax(f_logits, 1)
        expected = ((pi @ m.A) * q).sum(1, keepdim=True)
        cp, fp = m.get_type_probabilities(f_num, f_cat, c_num, c_cat)
    assert c_logits.shape == (B, 5) and f_logits.shape == (B, 5) and match.shape == (B, 1)
    assert torch.allclose(match, expected, atol=1e-5)
    assert torch.allclose(cp.sum(1), torch.ones(B, device=DEV), atol=1e-5)
    assert torch.allclose(fp.sum(1), torch.ones(B, device=DEV), atol=1e-5)
    assert "A" not in dict(m.named_parameters()) and "A" in dict(m.named_buffers()) and not m.A.requires_grad


@pytest.mark.parametrize("B", [2, 128, 16, 257, 4096])
def test_batches_vs_oracle(B):
    """config 2 shapes (train batches of 128, last val batch of 16) and tile edges of the 256-row head kernel."""
    # seeds checked for conditioning: no pre-activation within 1e-5 of a ReLU kink (a kink flips the unit's gradient)
    p = oracle.init_structural_params(12, META["firm_cat_cards"], 2, META["ceo_cat_cards"], seed=1000 + B)
    gen = torch.Generator().manual_seed(1000 + B)
    f_num, c_num = torch.randn(B, 12, generator=gen), torch.randn(B, 2, generator=gen)
    f_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in META["firm_cat_cards"]], 1)
    c_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in META["ceo_cat_cards"]], 1)
    tc = torch.distributions.Dirichlet(torch.ones(5)).sample((B,))
    tf = torch.distributions.Dirichlet(torch.ones(5)).sample((B,))
    m = _model(p)
    _zero_dropout(m)
    m.train()
    c_logits, f_logits, match = m(*[x.to(DEV) for x in (f_num, f_cat, c_num, c_cat)])
    loss = m.distillation_loss(c_logits, f_logits, tc.to(DEV), tf.to(DEV))
    loss.backward()
    po = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and k != "A") for k, v in p.items()}
    co, fo, mo = oracle.structural_forward(po, f_num, f_cat, c_num, c_cat, training=True)
    lo = oracle.structural_kl_loss(co, fo, tc, tf)
    lo.backward()
    assert_close_scaled(c_logits, co, TOL, "c_logits")
    assert_close_scaled(match, mo, TOL, "match")
    assert_close_scaled(loss, lo, TOL, "loss")
    check_grads(m, {k: po[k].grad for k, _ in m.named_parameters()}, 1e-4)
