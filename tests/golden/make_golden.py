"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference package cannot be imported whole here (its ``__init__`` pulls
matplotlib/shap/seaborn, which are absent), so its sub-modules are loaded under
the alias ``ref_cfm`` with the package ``__init__`` bypassed.  Dropout layers are
set to p=0 for the train-mode vectors (torch's Philox stream cannot be restated);
everything else is the reference's unmodified code path on CPU fp32.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    pkg = types.ModuleType("ref_cfm")
    pkg.__path__ = [os.path.join(REF, "ceo_firm_matching")]
    sys.modules["ref_cfm"] = pkg
    mods = {}
    for name in ["config", "model", "contrastive", "structural_config", "structural_model",
                 "structural_training", "training", "synthetic", "data", "structural_data"]:
        mods[name] = importlib.import_module(f"ref_cfm.{name}")
    return mods


def sd_np(model):
    return {"param/" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def grads_np(model):
    return {"grad/" + k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters()}


def zero_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0


def randomise_bn(model, gen):
    """Non-trivial BN affine + running stats so eval mode exercises them."""
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(1 + 0.3 * torch.randn(m.weight.shape, generator=gen))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=gen))
                m.running_mean.copy_(0.1 * torch.randn(m.bias.shape, generator=gen))
                m.running_var.copy_(0.5 + torch.rand(m.bias.shape, generator=gen))


def make_inputs(gen, B, n_fn, f_cards, n_cn, c_cards):
    f_num = torch.randn(B, n_fn, generator=gen)
    c_num = torch.randn(B, n_cn, generator=gen)
    f_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1)
    c_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1)
    return f_num, f_cat, c_num, c_cat


def golden_two_tower(mods, name, B, f_cards, c_cards, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    model = mods["model"].CEOFirmMatcher(meta, mods["config"].Config())
    zero_dropout(model)
    randomise_bn(model, gen)
    f_num, f_cat, c_num, c_cat = make_inputs(gen, B, 12, f_cards, 2, c_cards)
    target = torch.randn(B, 1, generator=gen)
    weights = 1.0 / (torch.rand(B, 1, generator=gen) * 0.9 + 0.1) ** 2
    out = {"f_num": f_num, "f_cat": f_cat, "c_num": c_num, "c_cat": c_cat, "target": target, "weights": weights}
    out = {k: v.numpy() for k, v in out.items()}
    out.update(sd_np(model))
    model.eval()
    with torch.no_grad():
        out["eval_score"] = model(f_num, f_cat, c_num, c_cat).numpy()
    model.train()
    preds = model(f_num, f_cat, c_num, c_cat)
    loss = (weights * (preds - target) ** 2).mean()          # training.py:52
    loss.backward()
    out["train_score"] = preds.detach().numpy()
    out["train_loss"] = loss.detach().numpy()
    out.update(grads_np(model))
    out.update({"after/" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()
                if "running" in k or "num_batches" in k})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss))


def golden_structural(mods, name, B, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    f_cards, c_cards = [4, 4, 2, 2], [2, 4, 2, 2, 2, 2, 2]
    meta = {"n_firm_num": 12, "n_ceo_num": 2, "firm_cat_cards": f_cards, "ceo_cat_cards": c_cards}
    cfg = mods["structural_config"].StructuralConfig()
    model = mods["structural_model"].StructuralDistillationNet(meta, cfg)
    zero_dropout(model)
    randomise_bn(model, gen)
    f_num, f_cat, c_num, c_cat = make_inputs(gen, B, 12, f_cards, 2, c_cards)
    t_ceo = torch.softmax(torch.randn(B, 5, generator=gen), 1)
    t_firm = torch.softmax(torch.randn(B, 5, generator=gen), 1)
    t_firm[0] = torch.tensor([0.0, 0.5, 0.5, 0.0, 0.0])      # xlogy(0, .) = 0 edge case
    out = {"f_num": f_num, "f_cat": f_cat, "c_num": c_num, "c_cat": c_cat, "target_ceo": t_ceo, "target_firm": t_firm}
    out = {k: v.numpy() for k, v in out.items()}
    out.update(sd_np(model))
    # eval forward + input sensitivities (structural_explain.py:68-89)
    model.eval()
    fn, cn = f_num.clone().requires_grad_(True), c_num.clone().requires_grad_(True)
    c_logits, f_logits, match = model(fn, f_cat, cn, c_cat)
    match.sum().backward()
    out["eval_c_logits"], out["eval_f_logits"] = c_logits.detach().numpy(), f_logits.detach().numpy()
    out["eval_match"] = match.detach().numpy()
    out["eval_dmatch_df_num"], out["eval_dmatch_dc_num"] = fn.grad.numpy().copy(), cn.grad.numpy().copy()
    model.zero_grad()
    # train step (structural_training.py:71-79)
    model.train()
    crit = torch.nn.KLDivLoss(reduction="batchmean")
    c_logits, f_logits, match = model(f_num, f_cat, c_num, c_cat)
    loss = crit(torch.log_softmax(c_logits, 1), t_ceo) + crit(torch.log_softmax(f_logits, 1), t_firm)
    loss.backward()
    out["train_c_logits"], out["train_f_logits"] = c_logits.detach().numpy(), f_logits.detach().numpy()
    out["train_match"], out["train_loss"] = match.detach().numpy(), loss.detach().numpy()
    out.update(grads_np(model))
    out.update({"after/" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()
                if "running" in k or "num_batches" in k})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss))


def golden_infonce(mods, name, B, D, seed):
    gen = torch.Generator().manual_seed(seed)
    f = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1).requires_grad_(True)
    c = torch.nn.functional.normalize(torch.randn(B, D, generator=gen), dim=1).requires_grad_(True)
    loss = mods["contrastive"].info_nce_loss(f, c, 0.07)
    loss.backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), firm_proj=f.detach().numpy(), ceo_proj=c.detach().numpy(),
                        loss=loss.detach().numpy(), d_firm=f.grad.numpy(), d_ceo=c.grad.numpy(), temperature=0.07)
    print(name, "loss", float(loss))


def golden_contrastive_model(mods, name, B, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    f_cards, c_cards = [5, 3, 2, 2], [2, 4, 2, 2, 2, 2, 3]
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    model = mods["contrastive"].ContrastiveCEOFirmMatcher(meta, mods["config"].Config())
    zero_dropout(model)
    randomise_bn(model, gen)
    f_num, f_cat, c_num, c_cat = make_inputs(gen, B, 12, f_cards, 2, c_cards)
    target = torch.randn(B, 1, generator=gen)
    weights = 1.0 / (torch.rand(B, 1, generator=gen) * 0.9 + 0.1) ** 2
    out = {"f_num": f_num, "f_cat": f_cat, "c_num": c_num, "c_cat": c_cat, "target": target, "weights": weights}
    out = {k: v.numpy() for k, v in out.items()}
    out.update(sd_np(model))
    model.train()
    score, fp, cp = model(f_num, f_cat, c_num, c_cat)
    mse = (weights * (score - target) ** 2).mean()            # contrastive.py:248
    cl = mods["contrastive"].info_nce_loss(fp, cp, 0.07)       # contrastive.py:254
    loss = 0.7 * mse + 0.3 * cl                                # contrastive.py:257
    loss.backward()
    out.update(train_score=score.detach().numpy(), firm_proj=fp.detach().numpy(), ceo_proj=cp.detach().numpy(),
               mse=mse.detach().numpy(), cl=cl.detach().numpy(), loss=loss.detach().numpy())
    out.update(grads_np(model))
    # retrieval metrics on eval embeddings (contrastive.py:275-332)
    cfg = mods["config"].Config()
    cfg.DEVICE = torch.device("cpu")
    data = {"firm_numeric": f_num, "firm_cat": f_cat, "ceo_numeric": c_num, "ceo_cat": c_cat}
    met = mods["contrastive"].compute_retrieval_metrics(model, data, cfg)
    model.eval()
    with torch.no_grad():
        fe, ce = model.get_embeddings(f_num, f_cat, c_num, c_cat)
    out.update(eval_firm_emb=fe.numpy(), eval_ceo_emb=ce.numpy())
    out.update({"metric/" + k: np.float64(v) for k, v in met.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss), met)


def golden_allpairs(name, F_, C_, D, seed):
    """analytical_extensions.py:471,483: torch.mm(firm, ceo.t()) * scale ; np.argsort(-scores)."""
    gen = torch.Generator().manual_seed(seed)
    u = torch.nn.functional.normalize(torch.randn(F_, D, generator=gen), dim=1)
    v = torch.nn.functional.normalize(torch.randn(C_, D, generator=gen), dim=1)
    scale = float(np.exp(np.log(1 / 0.07)))
    scores = (torch.mm(u, v.t()) * scale).numpy()
    ranking = np.stack([np.argsort(-scores[i]) for i in range(F_)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), u=u.numpy(), v=v.numpy(), scale=scale,
                        scores=scores, ranking=ranking)
    print(name, scores.shape)


def main():
    mods = load_reference()
    golden_two_tower(mods, "two_tower_b37", 37, [4, 4, 2, 2], [2, 4, 2, 2, 2, 2, 2], seed=11)
    golden_two_tower(mods, "two_tower_b300_bigcards", 300, [97, 31, 2, 7], [2, 4, 50, 2, 2, 13, 2], seed=12)
    golden_structural(mods, "structural_b29", 29, seed=21)
    golden_structural(mods, "structural_b200", 200, seed=22)
    golden_infonce(mods, "infonce_b33_d30", 33, 30, seed=31)
    golden_infonce(mods, "infonce_b200_d128", 200, 128, seed=32)
    golden_contrastive_model(mods, "contrastive_b64", 64, seed=41)
    golden_allpairs("allpairs_50x70_d60", 50, 70, 60, seed=51)


if __name__ == "__main__":
    main()
