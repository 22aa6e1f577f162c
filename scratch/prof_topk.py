import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching.scoring import score_topk
from ceo_firm_matching import ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
u = F.normalize(torch.randn(18944, 60, device=dev, generator=g), dim=1)   # 148 row blocks: exactly one wave
v = F.normalize(torch.randn(262144, 60, device=dev, generator=g), dim=1)
score_topk(u, v, 100, 14.2857)
f = F.normalize(torch.randn(18944, 128, device=dev, generator=g), dim=1); c = F.normalize(torch.randn(65536, 128, device=dev, generator=g), dim=1)
fb, cb = ops.pack_bf16(f), ops.pack_bf16(c)
rs_f, diag = ops.infonce_rowsum(fb, cb, 0.07, 0)
rs_c, _ = ops.infonce_rowsum(cb, fb, 0.07, 0, False)
ops.infonce_grad(fb, cb, 128, 0.07, 0, 65536, rs_f, rs_c, diag, torch.ones((), device=dev))
torch.cuda.synchronize(); print("done")
