import os, sys, torch
ROOT = '/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/ceo-recommender_b200'); sys.path.insert(0, ROOT + '/tests')
import oracle
from helpers import load_into
from ceo_firm_matching import CEOFirmMatcher, Config
dev = torch.device("cuda", 0)
torch.set_printoptions(linewidth=200, precision=4)
for cards0, seed, B, prec, key in ((5000, 3, 64, "fp32", "ceo_tower.5.bias"), (5000, 3, 300, "fp32", "firm_tower.5.bias"), (5000, 3, 64, "tf32", "firm_tower.1.bias")):
    f_cards, c_cards = [cards0, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2]
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=seed)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    gen = torch.Generator().manual_seed(100)
    sh = [torch.randn(B, 12, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1),
          torch.randn(B, 2, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1),
          torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5]
    model = load_into(CEOFirmMatcher(meta, Config()), p).to(dev).train()
    model.set_precision(prec)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
    loss, preds = model.forward_loss(*[t.to(dev) for t in sh])
    loss.backward()
    po = {k: (v.clone().double() if v.is_floating_point() else v.clone()) for k, v in p.items()}
    po = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in po.items()}
    ins = [t.double() if t.is_floating_point() else t for t in sh]
    lo = oracle.weighted_mse(oracle.two_tower_forward(po, *ins[:4], training=True), ins[4], ins[5]); lo.backward()
    g = dict(model.named_parameters())[key].grad.cpu().double(); e = po[key].grad
    bad = ((g - e).abs() > 1e-3 * e.abs().max()).nonzero().flatten()
    print(cards0, seed, B, prec, key, "n", g.numel(), "bad cols", bad.tolist())
    print("  got ", g[bad][:8]); print("  want", e[bad][:8]); print("  diff", (g - e)[bad][:8], " max|e|", float(e.abs().max()))
    kw = key.replace("bias", "weight")
    gw = dict(model.named_parameters())[kw].grad.cpu().double(); ew = po[kw].grad
    print("  gamma-grad at bad cols: got", gw[bad][:8], "want", ew[bad][:8])
