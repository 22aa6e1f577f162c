"""All-pairs top-100 at the config-5 column count (1M firms) for 131 072 CEOs (1/7.6 of config 5), best of 3."""
import sys
import torch
sys.path.insert(0, "ceo-recommender_b200")
import torch.nn.functional as F
from ceo_firm_matching.scoring import score_topk

g = torch.Generator(device="cuda").manual_seed(7)
ceos = F.normalize(torch.randn(131072, 60, device="cuda", generator=g), dim=1)
firms = F.normalize(torch.randn(1_000_000, 60, device="cuda", generator=g), dim=1)
score_topk(ceos[:4096], firms[:65536], 100, 1 / 0.07)
best = 1e9
for _ in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s, i, fl = score_topk(ceos, firms, 100, 1 / 0.07, return_flags=True); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("topk 131072 x 1M ms", best, "flagged rows", int(fl.sum()))
