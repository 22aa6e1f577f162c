import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200'); sys.path.insert(0, '/root/repo/tests')
import oracle
from helpers import load_into, dead_bias_names
from ceo_firm_matching import CEOFirmMatcher, Config
f_cards, c_cards, B = [50, 5, 3, 2], [2, 4, 3, 2, 2, 5, 2], 3000
for seed in (6, 7):
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=seed)
    g = torch.Generator().manual_seed(seed)
    ins = [torch.randn(B, 12, generator=g), torch.stack([torch.randint(0, n, (B,), generator=g) for n in f_cards], 1),
           torch.randn(B, 2, generator=g), torch.stack([torch.randint(0, n, (B,), generator=g) for n in c_cards], 1),
           torch.randn(B, 1, generator=g), 1.0 / (torch.rand(B, 1, generator=g) * 0.9 + 0.1) ** 2]
    po = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
    lo = oracle.weighted_mse(oracle.two_tower_forward(po, *ins[:4], training=True), ins[4], ins[5]); lo.backward()
    for prec in ("fp32", "tf32"):
        cfg = Config(); cfg.LATENT_DIM = 60
        meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
        m = load_into(CEOFirmMatcher(meta, cfg), p).cuda().train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout): mod.p = 0.0
        m.set_precision(prec)
        loss, preds = m.forward_loss(*[x.cuda() for x in ins]); loss.backward()
        dead = dead_bias_names(m)
        errs = {k: float((q.grad.cpu().double() - po[k].grad.double()).norm() / (po[k].grad.double().norm() + 1e-30))
                for k, q in m.named_parameters() if k not in dead}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        print(seed, prec, "loss rel", abs(float(loss) - float(lo)) / float(lo), [(k, f"{v:.2e}") for k, v in worst])
