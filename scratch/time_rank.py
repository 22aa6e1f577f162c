"""Retrieval ranks of 262 144 x 262 144 pairs (D = 60) through the tensor-core filter: trained-like positives and
untrained (random) embeddings, best of 2."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching.scoring import diagonal_ranks
g = torch.Generator(device="cuda").manual_seed(5)
N = 262144
f = F.normalize(torch.randn(N, 60, device="cuda", generator=g), dim=1)
r = F.normalize(torch.randn(N, 60, device="cuda", generator=g), dim=1)
c = F.normalize(0.8 * f + r, dim=1)
diagonal_ranks(f[:8192], c[:8192], method="tensor")
for name, cols, rows in (("trained-like", c, N), ("random (65536 rows)", r, 65536)):
    best = 1e9
    for _ in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = diagonal_ranks(f[:rows], cols, method="tensor"); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(name, "ms", best, "median rank", float(out.float().median()))
