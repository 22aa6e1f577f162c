"""Rank of the positive for 65 536 firms against 262 144 CEOs (D = 60) through the tensor-core filter, for an ncu
capture of simtile_kernel<SIM_RANK> (one launch)."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching.scoring import diagonal_ranks
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(3)
f = F.normalize(torch.randn(65536, 60, device=dev, generator=g), dim=1)
c = F.normalize(torch.randn(262144, 60, device=dev, generator=g), dim=1)
c[:65536] = F.normalize(0.8 * f + c[:65536], dim=1)
r = diagonal_ranks(f, c, method="tensor")
torch.cuda.synchronize()
print("ok median rank", float(r.float().median()))
