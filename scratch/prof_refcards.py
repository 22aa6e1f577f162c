"""Per-kernel-family device times of the config-4 batch shape with the reference's table cardinalities (2-4 classes)."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching import _native as N, CEOFirmMatcher, Config
from ceo_firm_matching.training import eager_step
dev = torch.device('cuda', 0)
lib = N.lib()
names = ["fwd1", "fwd2", "fwd3", "bwd1", "bwd2", "bwd3", "head", "emb", "reduce"]
f_cards, c_cards = [4, 4, 2, 2], [2, 4, 2, 2, 2, 2, 2]
meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
model = CEOFirmMatcher(meta, Config()).to(dev).train()
model.use_persistent_table_grads(True)
g = torch.Generator(device=dev).manual_seed(1)
B = bench.B_PER_GPU
base = bench.make_batches(4, B, dev, 1234)
batches = []
for b in base:
    f_cat = torch.stack([torch.randint(0, n, (B,), device=dev, generator=g) for n in f_cards], 1)
    c_cat = torch.stack([torch.randint(0, n, (B,), device=dev, generator=g) for n in c_cards], 1)
    batches.append((b[0], f_cat, b[2], c_cat, b[4], b[5]))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for i in range(3):
        eager_step(model, None, batches[i])
    torch.cuda.synchronize()
    lib.cfm_profile_enable(1)
    for i in range(10):
        eager_step(model, None, batches[i % 4])
    torch.cuda.synchronize()
ms = (C.c_double * 14)(); n = (C.c_int64 * 14)()
N.check(lib.cfm_profile_read(ms, n, 14)); lib.cfm_profile_enable(0)
print({k: round(ms[i] / 10, 4) for i, k in enumerate(names)}, "sum", round(sum(ms) / 10, 4))
