"""All-pairs top-100 at the config-5 column count (1M firms, D = 60) for a 75 776-row slice of the CEOs (2 row blocks x
148 SMs x 2 waves): the similarity kernel runs exactly as in the full 1M x 1M call, 13.5x shorter (for ncu)."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching.scoring import score_topk
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
u = F.normalize(torch.randn(75776, 60, device=dev, generator=g), dim=1)
v = F.normalize(torch.randn(1000000, 60, device=dev, generator=g), dim=1)
score_topk(u, v, 100, 14.2857)
torch.cuda.synchronize(); print("done")
