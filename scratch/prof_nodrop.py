import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching import _native as N
from ceo_firm_matching.training import eager_step
dev = torch.device('cuda', 0); lib = N.lib()
names = ["fwd1", "fwd2", "fwd3", "bwd1", "bwd2", "bwd3", "head", "emb", "reduce"]
for p in (0.1, 0.0):
    model = bench.build_model(dev, "fp32")
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout): m.p = p
    batches = bench.make_batches(8, bench.B_PER_GPU, dev, 1234)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(3): eager_step(model, None, batches[i])
        torch.cuda.synchronize(); lib.cfm_profile_enable(1)
        for i in range(10): eager_step(model, None, batches[i % 8])
        torch.cuda.synchronize()
    ms = (C.c_double * 14)(); n = (C.c_int64 * 14)()
    N.check(lib.cfm_profile_read(ms, n, 14)); lib.cfm_profile_enable(0)
    print("dropout", p, {k: round(ms[i] / 10, 4) for i, k in enumerate(names)})
    del model
