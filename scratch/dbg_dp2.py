import os, sys, torch
ROOT = '/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/ceo-recommender_b200'); sys.path.insert(0, ROOT + '/tests')
import oracle
from helpers import load_into, dead_bias_names
from ceo_firm_matching import CEOFirmMatcher, Config
dev = torch.device("cuda", 0)
for cards0, seed, B in ((5000, 3, 300), (5000, 4, 300), (500, 3, 300), (5000, 3, 1000), (5000, 3, 64), (5000, 3, 128)):
    f_cards, c_cards = [cards0, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2]
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=seed)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    gen = torch.Generator().manual_seed(100)
    sh = [torch.randn(B, 12, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1),
          torch.randn(B, 2, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1),
          torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5]
    model = load_into(CEOFirmMatcher(meta, Config()), p).to(dev).train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
    loss, preds = model.forward_loss(*[t.to(dev) for t in sh])
    loss.backward()
    res = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        po = {k: (v.clone().to(dt) if v.is_floating_point() else v.clone()) for k, v in p.items()}
        po = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in po.items()}
        ins = [t.to(dt) if t.is_floating_point() else t for t in sh]
        lo = oracle.weighted_mse(oracle.two_tower_forward(po, *ins[:4], training=True), ins[4], ins[5]); lo.backward()
        res[name] = {k: v.grad.double() for k, v in po.items() if v.grad is not None}
    dead = dead_bias_names(model)
    def worst(a, b):
        w = [(float((a[k] - b[k]).abs().max() / (b[k].abs().max() + 1e-30)), k) for k in a if k not in dead]
        w.sort(reverse=True); return [(f"{x:.1e}", k) for x, k in w[:3]]
    gpu = {k: q.grad.cpu().double() for k, q in model.named_parameters()}
    print(cards0, seed, B, "gpu-f64", worst(gpu, res["f64"]), "| f32-f64", worst(res["f32"], res["f64"]))
