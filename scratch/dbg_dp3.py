import os, sys, torch
ROOT = '/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/ceo-recommender_b200'); sys.path.insert(0, ROOT + '/tests')
import oracle
from helpers import load_into, dead_bias_names
from ceo_firm_matching import CEOFirmMatcher, Config
dev = torch.device("cuda", 0)
for cards0, seed, B, prec in ((5000, 3, 64, "fp32"), (5000, 3, 64, "tf32"), (5000, 3, 300, "fp32")):
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
    outs = []
    for rep in range(2):
        model.zero_grad()
        loss, preds = model.forward_loss(*[t.to(dev) for t in sh])
        loss.backward()
        outs.append({k: q.grad.clone() for k, q in model.named_parameters()})
    print("repeatable:", all(torch.equal(outs[0][k], outs[1][k]) for k in outs[0]))
    po = {k: (v.clone().double() if v.is_floating_point() else v.clone()) for k, v in p.items()}
    po = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in po.items()}
    ins = [t.double() if t.is_floating_point() else t for t in sh]
    # oracle with intermediates
    lo = oracle.weighted_mse(oracle.two_tower_forward(po, *ins[:4], training=True), ins[4], ins[5]); lo.backward()
    dead = dead_bias_names(model)
    print(cards0, seed, B, prec)
    for k, q in model.named_parameters():
        e = po[k].grad
        err = float((q.grad.cpu().double() - e).abs().max() / (e.abs().max() + 1e-30))
        flag = " dead" if k in dead else ""
        if err > 1e-4 or "ceo_tower" in k:
            print(f"   {k:28s} {err:.2e}{flag}")
    # hidden pre-activation kink census on the ceo tower
    x = oracle.gather_concat(po, "ceo", ins[2], ins[3]) if hasattr(oracle, "gather_concat") else None
