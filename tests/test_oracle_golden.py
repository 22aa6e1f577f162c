"""Pins the CPU oracle (oracle/) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py) and against the reference's one known-answer test."""
import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close, load_golden, params_from, t

# the oracle calls the same ATen ops as the reference in the same order -> near bit-exact
RT, AT = 1e-6, 1e-7


def _grads(params, loss):
    names = [k for k, v in params.items() if v.is_floating_point() and "running" not in k and k != "A"]
    for k in names:
        params[k].requires_grad_(True)
    return names


@pytest.mark.parametrize("name", ["two_tower_b37", "two_tower_b300_bigcards"])
def test_two_tower_matches_reference(name):
    g = load_golden(name)
    p = params_from(g)
    ins = [t(g[k]) for k in ("f_num", "f_cat", "c_num", "c_cat")]
    with torch.no_grad():
        assert_close(oracle.two_tower_forward(p, *ins, training=False), g["eval_score"], RT, AT, "eval score")
    names = _grads(p, None)
    preds = oracle.two_tower_forward(p, *ins, training=True)
    loss = oracle.weighted_mse(preds, t(g["target"]), t(g["weights"]))
    loss.backward()
    assert_close(preds, g["train_score"], RT, AT, "train score")
    assert_close(loss, g["train_loss"], RT, AT, "loss")
    for k in names:
        assert_close(p[k].grad, g["grad/" + k], 1e-5, 1e-6, "grad " + k)
    for k, v in g.items():
        if k.startswith("after/"):
            assert_close(p[k[6:]].detach(), v, RT, AT, k)


@pytest.mark.parametrize("name", ["structural_b29", "structural_b200"])
def test_structural_matches_reference(name):
    g = load_golden(name)
    p = params_from(g)
    f_num, f_cat, c_num, c_cat = [t(g[k]) for k in ("f_num", "f_cat", "c_num", "c_cat")]
    fn, cn = f_num.clone().requires_grad_(True), c_num.clone().requires_grad_(True)
    c_logits, f_logits, match = oracle.structural_forward(p, fn, f_cat, cn, c_cat, training=False)
    match.sum().backward()
    assert_close(c_logits, g["eval_c_logits"], RT, AT, "eval c_logits")
    assert_close(f_logits, g["eval_f_logits"], RT, AT, "eval f_logits")
    assert_close(match, g["eval_match"], RT, AT, "eval match")
    assert_close(fn.grad, g["eval_dmatch_df_num"], 1e-5, 1e-7, "dmatch/df_num")
    assert_close(cn.grad, g["eval_dmatch_dc_num"], 1e-5, 1e-7, "dmatch/dc_num")
    names = _grads(p, None)
    c_logits, f_logits, match = oracle.structural_forward(p, f_num, f_cat, c_num, c_cat, training=True)
    loss = oracle.structural_kl_loss(c_logits, f_logits, t(g["target_ceo"]), t(g["target_firm"]))
    loss.backward()
    assert_close(loss, g["train_loss"], RT, AT, "loss")
    assert_close(match, g["train_match"], RT, AT, "train match")
    for k in names:
        assert_close(p[k].grad, g["grad/" + k], 1e-5, 1e-6, "grad " + k)
    for k, v in g.items():
        if k.startswith("after/"):
            assert_close(p[k[6:]].detach(), v, RT, AT, k)


@pytest.mark.parametrize("name", ["infonce_b33_d30", "infonce_b200_d128"])
def test_infonce_matches_reference(name):
    g = load_golden(name)
    f, c = t(g["firm_proj"]).requires_grad_(True), t(g["ceo_proj"]).requires_grad_(True)
    loss = oracle.info_nce(f, c, float(g["temperature"]))
    loss.backward()
    assert_close(loss, g["loss"], RT, AT, "loss")
    assert_close(f.grad, g["d_firm"], 1e-5, 1e-7, "d_firm")
    assert_close(c.grad, g["d_ceo"], 1e-5, 1e-7, "d_ceo")


def test_infonce_b1_is_zero():
    # contrastive.py:124-126
    assert float(oracle.info_nce(torch.ones(1, 4), torch.ones(1, 4))) == 0.0


def test_contrastive_model_matches_reference():
    g = load_golden("contrastive_b64")
    p = params_from(g)
    ins = [t(g[k]) for k in ("f_num", "f_cat", "c_num", "c_cat")]
    names = [k for k, v in p.items() if v.is_floating_point() and "running" not in k]
    for k in names:
        p[k].requires_grad_(True)
    score, fp, cp = oracle.contrastive_forward(p, *ins, training=True)
    mse = oracle.weighted_mse(score, t(g["target"]), t(g["weights"]))
    cl = oracle.info_nce(fp, cp, 0.07)
    loss = 0.7 * mse + 0.3 * cl
    loss.backward()
    assert_close(score, g["train_score"], RT, 1e-6, "score")
    assert_close(fp, g["firm_proj"], RT, 1e-6, "firm_proj")
    assert_close(loss, g["loss"], RT, AT, "loss")
    for k in names:
        assert_close(p[k].grad, g["grad/" + k], 1e-5, 2e-6, "grad " + k)
    ranks = oracle.retrieval_ranks(t(g["eval_firm_emb"]), t(g["eval_ceo_emb"]))
    met = oracle.retrieval_metrics(ranks)
    for k, v in met.items():
        assert v == pytest.approx(float(g["metric/" + k]), abs=1e-12), k


def test_allpairs_matches_reference():
    g = load_golden("allpairs_50x70_d60")
    u, v, scale = t(g["u"]), t(g["v"]), float(g["scale"])
    assert_close(oracle.allpairs_scores(u, v, scale), g["scores"], RT, 1e-6, "scores")
    s, idx = oracle.allpairs_topk(u, v, 10, scale)
    np.testing.assert_array_equal(idx.numpy(), g["ranking"][:, :10])   # bit-exact index sets AND order
    assert_close(s, np.take_along_axis(g["scores"], g["ranking"][:, :10], 1), 1e-5, 1e-6, "topk scores")
    s_all, idx_all = oracle.allpairs_topk(u, v, 1000, scale)            # k > C clamps to C
    np.testing.assert_array_equal(idx_all.numpy(), g["ranking"])


def test_bilinear_known_answer():
    """Reference tests/test_structural_model.py:158-178: expected_match == sum((softmax(c) @ A) * softmax(f))."""
    p = oracle.init_structural_params(12, [4, 4, 2, 2], 2, [2, 4, 2, 2, 2, 2, 2], seed=3)
    gen = torch.Generator().manual_seed(5)
    B = 16
    f_num, c_num = torch.randn(B, 12, generator=gen), torch.randn(B, 2, generator=gen)
    f_cat = torch.randint(0, 2, (B, 4), generator=gen)
    c_cat = torch.randint(0, 2, (B, 7), generator=gen)
    with torch.no_grad():
        c_logits, f_logits, match = oracle.structural_forward(p, f_num, f_cat, c_num, c_cat)
    pi, q = torch.softmax(c_logits, 1), torch.softmax(f_logits, 1)
    expected = ((pi @ p["A"]) * q).sum(1, keepdim=True)
    assert torch.allclose(match, expected, atol=1e-5)
    assert torch.allclose(pi.sum(1), torch.ones(B), atol=1e-5)
    assert c_logits.shape == (B, 5) and f_logits.shape == (B, 5) and match.shape == (B, 1)


def test_bn_train_b1_raises():
    p = oracle.init_two_tower_params(12, [4, 4, 2, 2], 2, [2, 4, 2, 2, 2, 2, 2])
    with pytest.raises(ValueError):
        oracle.two_tower_forward(p, torch.randn(1, 12), torch.zeros(1, 4, dtype=torch.long),
                                 torch.randn(1, 2), torch.zeros(1, 7, dtype=torch.long), training=True)
