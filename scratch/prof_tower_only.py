"""Three eager two-tower steps at the config-4 shape (for ncu captures of the tower kernels)."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching.training import eager_step
dev = torch.device('cuda', 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
model = bench.build_model(dev, prec)
batches = bench.make_batches(2, bench.B_PER_GPU, dev, 1234)
for i in range(3):
    eager_step(model, None, batches[i % 2])
torch.cuda.synchronize()
print("ok")
