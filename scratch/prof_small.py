"""Short single-GPU run of every kernel family for ncu (launch list + full capture of the top kernels)."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching.training import eager_step
from ceo_firm_matching.contrastive import info_nce_loss
from ceo_firm_matching.scoring import score_topk
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
batches = bench.make_batches(3, bench.B_PER_GPU, dev, 1234)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for i in range(3):
        eager_step(model, None, batches[i])
torch.cuda.synchronize()
g = torch.Generator(device=dev).manual_seed(0)
f = F.normalize(torch.randn(16384, 128, device=dev, generator=g), dim=1).requires_grad_(True)
c = F.normalize(torch.randn(16384, 128, device=dev, generator=g), dim=1).requires_grad_(True)
info_nce_loss(f, c, 0.07).backward()
u = F.normalize(torch.randn(32768, 60, device=dev, generator=g), dim=1)
v = F.normalize(torch.randn(262144, 60, device=dev, generator=g), dim=1)
score_topk(u, v, 100, 14.2857)
torch.cuda.synchronize()
print("prof_small done")
