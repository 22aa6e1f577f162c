"""Dump the raw TMEM image of cfm_tc_selftest for tiny shapes (layout / descriptor debugging)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ceo-recommender_b200"))
import torch
from ceo_firm_matching import _native as Nn
torch.set_printoptions(precision=4, linewidth=200, sci_mode=False)
dev = torch.device("cuda", 0)

def run(mode, M, N, K, passes, ints=True):
    g = torch.Generator().manual_seed(1)
    sa = (K, M) if mode == 1 else (M, K)
    sb = (N, K) if mode == 0 else (K, N)
    if ints:
        A = torch.randint(-3, 4, sa, generator=g).float()
        B = torch.randint(-3, 4, sb, generator=g).float()
    else:
        A, B = torch.randn(*sa, generator=g), torch.randn(*sb, generator=g)
    cols = 32
    while cols < N: cols *= 2
    out = torch.full((128, cols), float("nan"), device=dev)
    Ad_, Bd_ = A.to(dev), B.to(dev)
    Nn.check(Nn.lib().cfm_tc_selftest(Nn.ptr(Ad_), Nn.ptr(Bd_), Nn.ptr(out), mode, M, N, K, passes, Nn.stream_ptr()))
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T) if mode == 0 else (A.double().T @ B.double() if mode == 1 else A.double() @ B.double())
    return out.cpu().double(), ref

for (mode, M, N, K, passes) in [(0, 128, 16, 8, 1), (0, 128, 16, 8, 3), (0, 128, 32, 32, 1), (0, 128, 64, 96, 1), (0, 128, 64, 96, 3),
                                 (2, 128, 32, 8, 1), (1, 64, 32, 8, 1), (1, 128, 32, 8, 1), (0, 64, 16, 8, 1)]:
    out, ref = run(mode, M, N, K, passes)
    print(f"=== mode {mode} M{M} N{N} K{K} passes {passes}: nan {int(out.isnan().sum())} zeros {int((out == 0).sum())} / {out.numel()}")
    print("ref[:4,:8]\n", ref[:4, :8])
    print("out[:4,:8]\n", out[:4, :8])
    ok = (out[:M, :N] - ref).abs().max() if M == 128 else None
    print("max err direct:", ok)
    if M == 64:
        print("out[16:20,:8]\n", out[16:20, :8]); print("out[32:36,:8]\n", out[32:36, :8]); print("ref[16:20,:8]\n", ref[16:20, :8])
