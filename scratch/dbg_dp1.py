import os, sys, torch
ROOT = '/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/ceo-recommender_b200'); sys.path.insert(0, ROOT + '/tests')
import oracle
from helpers import load_into, dead_bias_names
from ceo_firm_matching import CEOFirmMatcher, Config
dev = torch.device("cuda", 0)
for cards0 in (50, 5000):
  f_cards, c_cards, B = [cards0, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2], 300
  p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=3)
  meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
  gen = torch.Generator().manual_seed(100)
  sh = [torch.randn(B, 12, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1),
        torch.randn(B, 2, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1),
        torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5]
  for scale in (1.0, 0.5):
    for persistent in (False, True):
        model = load_into(CEOFirmMatcher(meta, Config()), p).to(dev).train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout): m.p = 0.0
        if persistent: model.use_persistent_table_grads(True)
        loss, preds = model.forward_loss(*[t.to(dev) for t in sh])
        (loss * scale).backward()
        po = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
        lo = oracle.weighted_mse(oracle.two_tower_forward(po, *sh[:4], training=True), sh[4], sh[5]); (lo * scale).backward()
        worst = []
        dead = dead_bias_names(model)
        for k, q in model.named_parameters():
            if k in dead: continue
            e = po[k].grad
            worst.append((float((q.grad.cpu() - e).abs().max() / (e.abs().max() + 1e-30)), k))
        worst.sort(reverse=True)
        print(cards0, scale, persistent, [(f"{a:.1e}", k) for a, k in worst[:4]])
