import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching.optim import FusedAdam
dev='cuda'
ps=[torch.randn(1_000_000, 48, device=dev).requires_grad_(True) for _ in range(4)]+[torch.randn(1_000_000, 8, device=dev).requires_grad_(True) for _ in range(7)]
for p in ps: p.grad=torch.randn_like(p)
n=sum(p.numel() for p in ps)
opt=FusedAdam(ps, lr=1e-3)
for _ in range(3): opt.step()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): opt.step()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print(f"adam step {ms:.4f} ms  {28*n/ms/1e6:.1f} GB/s")
