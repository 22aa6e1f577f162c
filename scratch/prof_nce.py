"""One InfoNCE forward+backward at config-3 size (B = 65 536, D = 128) for ncu captures of the similarity kernels:
launch 0 = row + column sums in one pass (SIM_ROWCOL), launch 1 = one gradient pass (SIM_GRAD)."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching import ops
dev = torch.device('cuda', 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator(device=dev).manual_seed(0)
f = F.normalize(torch.randn(B, 128, device=dev, generator=g), dim=1); c = F.normalize(torch.randn(B, 128, device=dev, generator=g), dim=1)
fb, cb = ops.pack_bf16(f), ops.pack_bf16(c)
rs_f, rs_c, diag = ops.infonce_rowcolsum(fb, cb, 0.07)
one = torch.ones((), device=dev)
d = ops.infonce_grad(fb, cb, 128, 0.07, 0, B, rs_f, rs_c, diag, one)
torch.cuda.synchronize()
print("ok", float(rs_f.sum()), float(rs_c.sum()), float(d.abs().sum()))
