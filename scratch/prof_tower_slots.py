"""Per-kernel-family device times of the config-4 step (eager re-issue, profile slots)."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching import _native as N
from ceo_firm_matching.training import eager_step, GraphedTwoTowerStep
dev = torch.device('cuda', 0)
lib = N.lib()
names = ["fwd1", "fwd2", "fwd3", "bwd1", "bwd2", "bwd3", "head", "emb", "reduce"]
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
model = bench.build_model(dev, prec)
batches = bench.make_batches(8, bench.B_PER_GPU, dev, 1234)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for i in range(3):
        eager_step(model, None, batches[i])
    torch.cuda.synchronize()
    lib.cfm_profile_enable(1)
    for i in range(10):
        eager_step(model, None, batches[i % 8])
    torch.cuda.synchronize()
ms = (C.c_double * 14)(); n = (C.c_int64 * 14)()
N.check(lib.cfm_profile_read(ms, n, 14)); lib.cfm_profile_enable(0)
print(prec, {k: round(ms[i] / 10, 4) for i, k in enumerate(names)}, "sum", round(sum(ms) / 10, 4))
model.zero_grad_fast()
runner = GraphedTwoTowerStep(model, batches[0], optimizer=None, warmup=3)
for i in range(5): runner.step(batches[i % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40): runner.step(batches[i % 8])
e1.record(); torch.cuda.synchronize()
print("graph step ms", e0.elapsed_time(e1) / 40)
