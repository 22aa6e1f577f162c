import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching.training import eager_step
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
batches = bench.make_batches(2, bench.B_PER_GPU, dev, 1234)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for i in range(2):
        eager_step(model, None, batches[i])
torch.cuda.synchronize(); print("done")
