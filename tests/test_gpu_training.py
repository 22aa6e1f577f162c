"""GPU tests of the training loops (drop-in ``train_model`` / ``train_structural_model`` / CLIs): graph-captured
steps equal eager steps, and whole short trainings follow the CPU oracle's trajectory (same init, same batches,
same Adam) when dropout is off."""
import contextlib
import io
import re

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

import oracle
from helpers import assert_close_scaled, check_grads, dead_bias_names, load_into

pytestmark = pytest.mark.gpu
DEV = "cuda"
F_CARDS, C_CARDS = [7, 5, 3, 2], [2, 4, 3, 2, 2, 5, 2]


def _batch(B, seed):
    gen = torch.Generator().manual_seed(seed)
    f_num, c_num = torch.randn(B, 12, generator=gen), torch.randn(B, 2, generator=gen)
    f_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in F_CARDS], 1)
    c_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in C_CARDS], 1)
    return [f_num, f_cat, c_num, c_cat, torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5]


def _model(p, dropout_off=True):
    from ceo_firm_matching import CEOFirmMatcher, Config
    meta = {"n_firm_numeric": 12, "firm_cat_counts": F_CARDS, "n_ceo_numeric": 2, "ceo_cat_counts": C_CARDS}
    m = load_into(CEOFirmMatcher(meta, Config()), p).to(DEV).train()
    if dropout_off:
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    return m


def test_graphed_step_equals_eager_step():
    from ceo_firm_matching.training import GraphedTwoTowerStep, eager_step
    p = oracle.init_two_tower_params(12, F_CARDS, 2, C_CARDS, seed=3)
    b0, b1 = [[t.to(DEV) for t in _batch(512, s)] for s in (1, 2)]
    eager = _model(p)
    eager.use_persistent_table_grads(True)
    graphed = _model(p)
    graphed.use_persistent_table_grads(True)
    runner = GraphedTwoTowerStep(graphed, b0, optimizer=None, warmup=2)
    load_into(graphed, {k: v.to(DEV) for k, v in p.items()})       # undo the BN running-stat updates of the warm-up
    for b in (b0, b1, b0):
        le = eager_step(eager, None, b)
        lg = runner.step(b)
        assert torch.equal(le, lg)                                   # same kernels, same order -> bitwise equal
        for (k, pe), (_, pg) in zip(eager.named_parameters(), graphed.named_parameters()):
            assert torch.equal(pe.grad, pg.grad), k
    for (k, be), (_, bg) in zip(eager.named_buffers(), graphed.named_buffers()):
        assert torch.equal(be, bg), k


def test_graphed_step_draws_fresh_dropout_masks():
    from ceo_firm_matching.training import GraphedTwoTowerStep
    p = oracle.init_two_tower_params(12, F_CARDS, 2, C_CARDS, seed=3)
    b = [t.to(DEV) for t in _batch(512, 1)]
    m = _model(p, dropout_off=False)
    m.use_persistent_table_grads(True)
    runner = GraphedTwoTowerStep(m, b, optimizer=None, warmup=1)
    losses = {float(runner.step(b)) for _ in range(4)}
    assert len(losses) == 4


class _DictDataset(torch.utils.data.Dataset):
    def __init__(self, data):
        self.data = data

    def __len__(self):
        return len(self.data["target"])

    def __getitem__(self, i):
        return {k: v[i] for k, v in self.data.items() if isinstance(v, torch.Tensor)}


def test_train_model_follows_oracle_trajectory(monkeypatch):
    """cli --synthetic shape (800 rows, batch 256 -> 256,256,256,32) for 3 epochs, dropout off, no shuffle:
    the printed epoch-0 loss and the final parameters must match an oracle loop (CPU autograd + torch Adam)."""
    import ceo_firm_matching as cfm
    from ceo_firm_matching import model as model_mod
    real_mlp = model_mod._mlp

    def mlp_no_dropout(i, o):
        seq = real_mlp(i, o)
        for mod in seq:
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        return seq

    monkeypatch.setattr(model_mod, "_mlp", mlp_no_dropout)
    cfg = cfm.Config()
    cfg.EPOCHS, cfg.DEVICE = 3, torch.device(DEV)
    proc = cfm.DataProcessor(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        df = proc.prepare_features(cfm.generate_synthetic_data(1000)).iloc[:800]
        proc.fit(df)
        data = proc.transform(df)
    loader = DataLoader(cfm.CEOFirmDataset(data), batch_size=256, shuffle=False)

    torch.manual_seed(7)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        model = cfm.train_model(loader, loader, data, cfg)
    assert next(model.parameters()).device.type == "cuda"
    printed = float(re.search(r"Epoch 0: Avg Train Loss = ([0-9.]+)", out.getvalue()).group(1))

    # oracle loop with identical init (same module construction order under the same seed)
    torch.manual_seed(7)
    init = cfm.CEOFirmMatcher(data, cfg).state_dict()
    p = {k: v.clone() for k, v in init.items()}
    names = [k for k, v in p.items() if v.is_floating_point() and "running" not in k]
    for k in names:
        p[k].requires_grad_(True)
    opt = torch.optim.Adam([p[k] for k in names], lr=cfg.LEARNING_RATE)
    keys = ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat", "target", "weights")
    first_epoch = None
    for epoch in range(3):
        tot = 0.0
        for s in range(0, 800, 256):
            f_num, f_cat, c_num, c_cat, tgt, w = [data[k][s:s + 256] for k in keys]
            opt.zero_grad()
            loss = oracle.weighted_mse(oracle.two_tower_forward(p, f_num, f_cat, c_num, c_cat, training=True), tgt, w)
            loss.backward()
            opt.step()
            tot += float(loss)
        first_epoch = tot / 4 if first_epoch is None else first_epoch
    assert printed == pytest.approx(first_epoch, rel=2e-4, abs=1e-4)
    sd = model.state_dict()
    # Biases that feed a train-mode BatchNorm have a mathematically zero gradient; Adam turns the rounding noise
    # of either implementation into +-lr random walks, and BN cancels them again: not comparable, not relevant.
    dead = dead_bias_names(model.train())
    for k in names:
        if k not in dead:
            assert_close_scaled(sd[k], p[k].detach(), 2e-3, "trained " + k, floor=1e-5)
    assert int(sd["firm_tower.1.num_batches_tracked"]) == 12


def test_cli_synthetic_runs_and_learns(capsys):
    from ceo_firm_matching import cli
    torch.manual_seed(0)
    model = cli.main(["--synthetic"])
    text = capsys.readouterr().out
    losses = [float(x) for x in re.findall(r"Avg Train Loss = ([0-9.]+)", text)]
    assert len(losses) == 8 and losses[-1] < losses[0]               # prints every 5th of 40 epochs
    assert "Validation weighted MSE" in text and model is not None


def test_train_structural_model_follows_oracle_trajectory():
    import ceo_firm_matching as cfm
    cfg = cfm.StructuralConfig()
    cfg.DATA_PATH, cfg.EPOCHS, cfg.BATCH_SIZE, cfg.DROPOUT, cfg.DEVICE = "SYNTHETIC_MODE", 2, 128, 0.0, torch.device(DEV)
    proc = cfm.StructuralDataProcessor(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        train_ds, val_ds, _ = proc.load_and_prep()
    meta = proc.get_metadata()
    train_loader = DataLoader(train_ds, batch_size=128, shuffle=False, drop_last=True)
    val_loader = DataLoader(val_ds, batch_size=128, shuffle=False)
    torch.manual_seed(11)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        model = cfm.train_structural_model(train_loader, val_loader, meta, cfg)
    got = re.search(r"Epoch\s+1: Train Loss = ([0-9.]+), Val Loss = ([0-9.]+)", out.getvalue())
    torch.manual_seed(11)
    p = {k: v.clone() for k, v in cfm.StructuralDistillationNet(meta, cfg).state_dict().items()}
    names = [k for k, v in p.items() if v.is_floating_point() and "running" not in k and k != "A"]
    for k in names:
        p[k].requires_grad_(True)
    opt = torch.optim.Adam([p[k] for k in names], lr=cfg.LEARNING_RATE)
    keys = ("firm_num", "firm_cat", "ceo_num", "ceo_cat", "target_ceo", "target_firm")

    def loss_of(d, s, e, training):
        f_num, f_cat, c_num, c_cat, tc, tf = [d[k][s:e] for k in keys]
        c, f, _ = oracle.structural_forward(p, f_num, f_cat, c_num, c_cat, training=training, dropout=0.0)
        return oracle.structural_kl_loss(c, f, tc, tf)

    for epoch in range(2):
        tr = 0.0
        for s in range(0, 1536, 128):                                 # 12 full batches (drop_last)
            opt.zero_grad()
            loss = loss_of(train_ds.data, s, s + 128, True)
            loss.backward()
            opt.step()
            tr += float(loss)
        with torch.no_grad():
            va = sum(float(loss_of(val_ds.data, s, min(s + 128, 400), False)) for s in range(0, 400, 128))
    assert float(got.group(1)) == pytest.approx(tr / 12, rel=5e-4, abs=1e-4)
    assert float(got.group(2)) == pytest.approx(va / 4, rel=5e-4, abs=1e-4)
    assert next(model.parameters()).device.type == "cuda"


def test_structural_cli_synthetic_runs(capsys):
    from ceo_firm_matching import structural_cli
    assert structural_cli.main(["--synthetic", "--epochs", "3", "--batch-size", "128"]) == 0
    text = capsys.readouterr().out
    assert "Training Complete" in text and "Train size: 1600, Val size: 400" in text
