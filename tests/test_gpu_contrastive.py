"""GPU parity of the product ``ContrastiveCEOFirmMatcher`` (forward, combined loss, every gradient) and of
``compute_retrieval_metrics`` against the golden vectors produced by the real reference (contrastive.py:52-99,
197-272, 275-332; fixture tests/golden/contrastive_b64.npz, generator tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import assert_close_scaled, check_grads, load_golden, load_into, params_from, t

pytestmark = pytest.mark.gpu
DEV = "cuda"
META = {"n_firm_numeric": 12, "firm_cat_counts": [5, 3, 2, 2], "n_ceo_numeric": 2, "ceo_cat_counts": [2, 4, 2, 2, 2, 2, 3]}


def _model(g):
    from ceo_firm_matching import Config
    from ceo_firm_matching.contrastive import ContrastiveCEOFirmMatcher
    m = ContrastiveCEOFirmMatcher(META, Config())
    load_into(m, params_from(g))
    return m.to(DEV)


def _after_one_train_forward(g):
    """The golden eval-mode outputs were taken after the training forward of the same script had updated the
    BatchNorm running statistics once (tests/golden/make_golden.py:golden_contrastive_model)."""
    m = _model(g).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    with torch.no_grad():
        m(*[t(g[k]).to(DEV) for k in ("f_num", "f_cat", "c_num", "c_cat")])
    return m


def test_state_dict_layout_is_the_reference_one():
    g = load_golden("contrastive_b64")
    m = _model(g)
    assert set(m.state_dict().keys()) == set(params_from(g).keys())


def test_train_forward_loss_and_all_gradients_match_reference():
    """One training step of train_contrastive's objective, dropout off (the golden run's setting):
    0.7 * weighted MSE + 0.3 * InfoNCE.  fp32-class towers and heads; the InfoNCE matrix runs in bf16 on the
    tensor cores, so every gradient carries the bf16 bar (4e-3 of the tensor's scale) and the scores the fp32 bar."""
    from ceo_firm_matching.contrastive import info_nce_loss
    g = load_golden("contrastive_b64")
    m = _model(g).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ins = [t(g[k]).to(DEV) for k in ("f_num", "f_cat", "c_num", "c_cat")]
    target, weights = t(g["target"]).to(DEV), t(g["weights"]).to(DEV)
    score, firm_proj, ceo_proj = m(*ins)
    assert score.shape == (64, 1) and firm_proj.shape == (64, 30) and ceo_proj.shape == (64, 30)
    mse = (weights * (score - target) ** 2).mean()
    cl = info_nce_loss(firm_proj, ceo_proj, 0.07)
    loss = 0.7 * mse + 0.3 * cl
    loss.backward()
    assert_close_scaled(score, g["train_score"], 2e-5, "train score")
    assert_close_scaled(firm_proj, g["firm_proj"], 2e-5, "firm_proj")
    assert_close_scaled(ceo_proj, g["ceo_proj"], 2e-5, "ceo_proj")
    assert float(mse) == pytest.approx(float(g["mse"]), rel=2e-5)
    assert float(cl) == pytest.approx(float(g["cl"]), rel=1e-3)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-3)
    # Towers: their gradients mix the exact MSE part with 0.3 x the InfoNCE part -> the bf16 bar of the InfoNCE tests
    # (4e-3 of the tensor's scale).  Projection heads: ALL of their gradient flows through the bf16 similarity
    # matrix, where rounding the operands to 2^-9 moves a logit s/T by up to 2^-9/0.07 = 2.8 % before the
    # exponential; the batch sum of such terms is held to 1.5e-2 of the tensor's scale (measured worst case 1.0e-2).
    expected = {k[5:]: v for k, v in g.items() if k.startswith("grad/")}
    heads = {k: v for k, v in expected.items() if "projector" in k}
    towers = {k: v for k, v in expected.items() if "projector" not in k}

    class _View:                                     # check_grads walks named_parameters(): restrict it to a subset
        def __init__(self, mod, keep):
            self.mod, self.keep = mod, keep

        def named_parameters(self):
            return [(k, q) for k, q in self.mod.named_parameters() if k in self.keep]

        def named_modules(self):
            return self.mod.named_modules()

    check_grads(_View(m, towers), towers, 4e-3, "contrastive towers")
    check_grads(_View(m, heads), heads, 1.5e-2, "contrastive heads")


def test_get_embeddings_eval_mode_matches_reference():
    g = load_golden("contrastive_b64")
    m = _after_one_train_forward(g).eval()
    ins = [t(g[k]).to(DEV) for k in ("f_num", "f_cat", "c_num", "c_cat")]
    with torch.no_grad():
        fe, ce = m.get_embeddings(*ins)
    assert_close_scaled(fe, g["eval_firm_emb"], 2e-5, "eval firm embeddings")
    assert_close_scaled(ce, g["eval_ceo_emb"], 2e-5, "eval ceo embeddings")
    assert torch.allclose(fe.norm(dim=1), torch.ones(64, device=DEV), atol=1e-5)


def test_compute_retrieval_metrics_matches_reference():
    """The product function end to end (eval forward -> rank of the diagonal counted on the device)."""
    from ceo_firm_matching import Config
    from ceo_firm_matching.contrastive import compute_retrieval_metrics
    g = load_golden("contrastive_b64")
    m = _after_one_train_forward(g)
    data = {"firm_numeric": t(g["f_num"]), "firm_cat": t(g["f_cat"]), "ceo_numeric": t(g["c_num"]), "ceo_cat": t(g["c_cat"])}
    cfg = Config()
    cfg.DEVICE = DEV
    met = compute_retrieval_metrics(m, data, cfg)
    assert set(met) == {"recall@1", "recall@5", "recall@10", "MRR", "median_rank"}
    for k, v in met.items():
        assert v == pytest.approx(float(g["metric/" + k]), abs=1e-9), k


def test_diagonal_ranks_match_a_full_sort():
    """rank_i = position of s_ii in the descending sort of row i (contrastive.py:310-320), ties broken by index."""
    from ceo_firm_matching.scoring import diagonal_ranks
    gen = torch.Generator().manual_seed(5)
    u = torch.nn.functional.normalize(torch.randn(700, 60, generator=gen), dim=1)
    v = torch.nn.functional.normalize(0.5 * u + torch.randn(700, 60, generator=gen), dim=1)
    ranks = diagonal_ranks(u.to(DEV), v.to(DEV)).cpu().numpy()
    sim = (u.double() @ v.double().t()).numpy()
    order = np.argsort(-sim, axis=1, kind="stable")
    want = np.array([int(np.where(order[i] == i)[0][0]) + 1 for i in range(700)])
    np.testing.assert_array_equal(ranks, want)


def test_train_contrastive_follows_oracle_trajectory(monkeypatch, capsys):
    """The whole loop of contrastive.py:197-272 (InfoNCE branch): 6 epochs over 512 synthetic rows in batches of 256,
    dropout off, no shuffle; the printed losses of epochs 0 and 5 and the parameter movement must follow an oracle loop
    (CPU autograd on oracle.contrastive_forward + oracle.info_nce, torch Adam) started from the same initial
    parameters.  The InfoNCE part runs in bf16 on the tensor cores: printed losses 2e-3; Adam's normalised update turns
    a flipped sign of a near-zero gradient entry into a +-lr step, so the movement of every tensor (trained minus
    initial) is compared in relative L2 norm (30 %)."""
    import contextlib
    import io
    import re
    from torch.utils.data import DataLoader
    import ceo_firm_matching as cfm
    import oracle
    from ceo_firm_matching import contrastive as cmod
    from ceo_firm_matching import model as model_mod
    from helpers import dead_bias_names
    real_mlp = model_mod._mlp

    def mlp_no_dropout(i, o):
        seq = real_mlp(i, o)
        for mod in seq:
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return seq

    monkeypatch.setattr(model_mod, "_mlp", mlp_no_dropout)
    cfg = cfm.Config()
    cfg.EPOCHS, cfg.DEVICE = 6, torch.device(DEV)
    proc = cfm.DataProcessor(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        df = proc.prepare_features(cfm.generate_synthetic_data(700)).iloc[:512]
        proc.fit(df)
        data = proc.transform(df)
    loader = DataLoader(cfm.CEOFirmDataset(data), batch_size=256, shuffle=False)

    torch.manual_seed(11)
    model = cmod.train_contrastive(loader, loader, data, cfg, contrastive_weight=0.3, temperature=0.07)
    printed = capsys.readouterr().out
    m0 = re.search(r"Epoch 0: Loss=([0-9.]+) \(MSE=([0-9.]+), CL=([0-9.]+)\)", printed)
    m5 = re.search(r"Epoch 5: Loss=([0-9.]+) \(MSE=([0-9.]+), CL=([0-9.]+)\)", printed)
    assert m0 and m5, printed
    assert "Loss type: InfoNCE" in printed and next(model.parameters()).device.type == "cuda"

    torch.manual_seed(11)
    init = cmod.ContrastiveCEOFirmMatcher(data, cfg).state_dict()
    p = {k: v.clone() for k, v in init.items()}
    names = [k for k, v in p.items() if v.is_floating_point() and "running" not in k]
    for k in names:
        p[k].requires_grad_(True)
    opt = torch.optim.Adam([p[k] for k in names], lr=cfg.LEARNING_RATE)
    keys = ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat", "target", "weights")
    per_epoch = []
    for epoch in range(6):
        tot = np.zeros(3)
        for s in range(0, 512, 256):
            f_num, f_cat, c_num, c_cat, tgt, w = [data[k][s:s + 256] for k in keys]
            opt.zero_grad()
            score, fp, cp = oracle.contrastive_forward(p, f_num, f_cat, c_num, c_cat, training=True)
            mse = oracle.weighted_mse(score, tgt, w)
            cl = oracle.info_nce(fp, cp, 0.07)
            loss = 0.7 * mse + 0.3 * cl
            loss.backward()
            opt.step()
            tot += np.array([float(loss), float(mse), float(cl)])
        per_epoch.append(tot / 2)
    for m, want3 in ((m0, per_epoch[0]), (m5, per_epoch[5])):
        for got, want in zip(m.groups(), want3):
            assert float(got) == pytest.approx(want, rel=2e-3, abs=2e-4)
    assert per_epoch[5][0] < per_epoch[0][0]                       # and the objective went down
    sd = model.state_dict()
    dead = {"base_model." + k for k in dead_bias_names(model.base_model.train())}
    for k in names:
        if k in dead:
            continue
        moved_o = (p[k].detach() - init[k]).double()
        moved_g = (sd[k].cpu() - init[k]).double()
        assert float((moved_g - moved_o).norm()) <= 0.3 * float(moved_o.norm()) + 1e-6, k
