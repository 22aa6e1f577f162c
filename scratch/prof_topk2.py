import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching.scoring import score_topk
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
u = F.normalize(torch.randn(75776, 60, device=dev, generator=g), dim=1)
v = F.normalize(torch.randn(524288, 60, device=dev, generator=g), dim=1)
score_topk(u, v, 100, 14.2857)
torch.cuda.synchronize(); print("done")
