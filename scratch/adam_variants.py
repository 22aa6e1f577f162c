"""Which contraction pattern of the fused Adam kernel is bit-equal to torch.optim.Adam(capturable=True, foreach)?"""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching.optim import FusedAdam
dev = 'cuda'
shapes = [(1000, 48), (7,), (64, 205), (1,), (333, 8), (70001,)]
for variant in range(4):
    g = torch.Generator(device=dev).manual_seed(0)
    pa = [torch.randn(s, device=dev, generator=g).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = torch.optim.Adam(pa, lr=1e-3, capturable=True, foreach=True)
    ob = FusedAdam(pb, lr=1e-3, variant=variant)
    ok = True; worst = 0.0
    for it in range(6):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device=dev, generator=g) * (10.0 ** (it - 3))
            if it == 4: gr[::2] = 0
            a.grad = gr.clone(); b.grad = gr.clone()
        oa.step(); ob.step()
        for a, b in zip(pa, pb):
            ok = ok and torch.equal(a, b)
            worst = max(worst, float(((a - b).abs() / (a.abs() + 1e-12)).max()))
    sa, sb = oa.state[pa[0]], ob.state[pb[0]]
    print("variant", variant, "bitwise", ok, "worst rel", worst, "step", float(sa["step"]), float(sb["step"]),
          "m equal", torch.equal(sa["exp_avg"], sb["exp_avg"]), "v equal", torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"]))
