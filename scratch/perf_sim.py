import sys, ctypes as C, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching import ops, _native as N
from ceo_firm_matching.scoring import score_topk
lib = N.lib(); dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
def prof(label, fn, reps=3):
    fn(); torch.cuda.synchronize()
    lib.cfm_profile_enable(1)
    for _ in range(reps): fn()
    ms = (C.c_double * 14)(); n = (C.c_int64 * 14)()
    N.check(lib.cfm_profile_read(ms, n, 14)); lib.cfm_profile_enable(0)
    names = ["fwd1","fwd2","fwd3","bwd1","bwd2","bwd3","head","emb","reduce","nce_rowsum","nce_grad","topk","topk_post"]
    print(label, {names[i]: round(ms[i]/n[i], 3) for i in range(13) if n[i]})
B, D = 65536, 128
f = F.normalize(torch.randn(B, D, device=dev, generator=g), dim=1); c = F.normalize(torch.randn(B, D, device=dev, generator=g), dim=1)
fb, cb = ops.pack_bf16(f), ops.pack_bf16(c)
rs_f, diag = ops.infonce_rowsum(fb, cb, 0.07); rs_c, _ = ops.infonce_rowsum(cb, fb, 0.07, want_diag=False)
one = torch.ones((), device=dev)
prof("infonce B=65536 D=128 (per call, ms)", lambda: (ops.infonce_rowsum(fb, cb, 0.07), ops.infonce_grad(fb, cb, D, 0.07, 0, B, rs_f, rs_c, diag, one)))
for R, Cn in [(131072, 1000000), (16384, 1000000)]:
    u = F.normalize(torch.randn(R, 60, device=dev, generator=g), dim=1); v = F.normalize(torch.randn(Cn, 60, device=dev, generator=g), dim=1)
    prof(f"topk {R}x{Cn} D=60 k=100 (ms)", lambda: score_topk(u, v, 100, 14.2857), reps=2)
    del u, v
