"""Device time of the InfoNCE forward building blocks at config-3 size: two row-sum passes against one row+column pass."""
import sys
import torch
sys.path.insert(0, "ceo-recommender_b200")
import torch.nn.functional as F
from ceo_firm_matching import ops

g = torch.Generator(device="cuda").manual_seed(0)
f = F.normalize(torch.randn(65536, 128, device="cuda", generator=g), dim=1)
c = F.normalize(torch.randn(65536, 128, device="cuda", generator=g), dim=1)
fb, cb = ops.pack_bf16(f), ops.pack_bf16(c)


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("rowsum one direction ms", t(lambda: ops.infonce_rowsum(fb, cb, 0.07)))
print("rowcolsum ms", t(lambda: ops.infonce_rowcolsum(fb, cb, 0.07)))
print("pack ms", t(lambda: ops.pack_bf16(f)))
