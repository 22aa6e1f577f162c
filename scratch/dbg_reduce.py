import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200'); sys.path.insert(0, '/root/repo/tests')
import oracle
from helpers import load_into
from ceo_firm_matching import CEOFirmMatcher, Config, ops
f_cards, c_cards = [5000, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2]
p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=3)
meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
m = load_into(CEOFirmMatcher(meta, Config()), p).cuda().train()
m.use_persistent_table_grads(True)
for B in (300, 600, 257):
    for hi, h in enumerate(m._handles):
        cards = f_cards if hi == 0 else c_cards
        g = torch.Generator().manual_seed(B + hi)
        x = torch.stack([torch.randint(0, n, (B,), generator=g) for n in cards], 1).cuda()
        dx = torch.randn(B, h.n_tables * h.emb_dim, generator=g).cuda()
        ops.reduce_table_grads(h, x, dx)
        for k, e in enumerate(h.embeddings):
            ref = torch.zeros_like(e.weight).index_add_(0, x[:, k], dx[:, k * h.emb_dim:(k + 1) * h.emb_dim])
            err = float((e.weight.grad - ref).abs().max())
            print(B, hi, k, "maxerr", err, "nnz rows", int((e.weight.grad.abs().sum(1) > 0).sum()), int((ref.abs().sum(1) > 0).sum()))
