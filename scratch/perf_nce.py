import sys, ctypes as C, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching import ops, _native as N
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
for mask in [int(a) for a in sys.argv[1:]] or [3]:
    lib.cfm_simtile_set_poly(mask)
    prof(f"infonce B=65536 D=128 poly={mask} (ms)", lambda: (ops.infonce_rowsum(fb, cb, 0.07), ops.infonce_grad(fb, cb, D, 0.07, 0, B, rs_f, rs_c, diag, one)))
