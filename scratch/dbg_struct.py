import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200'); sys.path.insert(0, '/root/repo/tests')
import oracle
from helpers import load_into
from ceo_firm_matching import StructuralConfig, StructuralDistillationNet
META = {"n_firm_num": 12, "n_ceo_num": 2, "firm_cat_cards": [4, 4, 2, 2], "ceo_cat_cards": [2, 4, 2, 2, 2, 2, 2]}
for B in (64, 128, 128, 192, 256, 256, 320, 128, 64, 512, 128):
    p = oracle.init_structural_params(12, META["firm_cat_cards"], 2, META["ceo_cat_cards"], seed=B)
    gen = torch.Generator().manual_seed(B)
    f_num, c_num = torch.randn(B, 12, generator=gen), torch.randn(B, 2, generator=gen)
    f_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in META["firm_cat_cards"]], 1)
    c_cat = torch.stack([torch.randint(0, n, (B,), generator=gen) for n in META["ceo_cat_cards"]], 1)
    tc = torch.distributions.Dirichlet(torch.ones(5)).sample((B,)); tf = torch.distributions.Dirichlet(torch.ones(5)).sample((B,))
    m = load_into(StructuralDistillationNet(META, StructuralConfig()), p).cuda().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout): mod.p = 0.0
    c_logits, f_logits, match = m(*[x.cuda() for x in (f_num, f_cat, c_num, c_cat)])
    loss = m.distillation_loss(c_logits, f_logits, tc.cuda(), tf.cuda()); loss.backward()
    po = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and k != "A") for k, v in p.items()}
    co, fo, mo = oracle.structural_forward(po, f_num, f_cat, c_num, c_cat, training=True)
    lo = oracle.structural_kl_loss(co, fo, tc, tf); lo.backward()
    worst = sorted(((float((prm.grad.cpu() - po[k].grad).abs().max() / (po[k].grad.abs().max() + 1e-12)), k) for k, prm in m.named_parameters()), reverse=True)[:4]
    print(B, [(round(e,5), k) for e, k in worst if "0.bias" not in k][:2])
