"""globaltimer trace of one tcgen05 tower stage launch (config-4 shape).  argv[1] = 10*(0 fwd,1 bwd)+stage."""
import sys, torch, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching import _native as N
from ceo_firm_matching.training import eager_step
dev = torch.device('cuda', 0)
model = bench.build_model(dev, "fp32")
batches = bench.make_batches(2, bench.B_PER_GPU, dev, 1234)
for i in range(3):
    eager_step(model, None, batches[i % 2])
torch.cuda.synchronize()
lib = N.lib()
for code in [int(a) for a in sys.argv[1:]] or [1]:
    buf = torch.zeros(2 * 148 * 128, dtype=torch.int64, device=dev)
    N.check(lib.cfm_debug_set_trace(N.ptr(buf), code))
    eager_step(model, None, batches[0])
    torch.cuda.synchronize()
    N.check(lib.cfm_debug_set_trace(None, 0))
    t = buf.cpu().view(2, 148, 128)
    print("=== code", code)
    for tower in (0, 1):
        for cta in (0, 73, 147):
            r = t[tower, cta]
            t0 = int(r[0])
            def rel(x): return round((int(x) - t0) / 1000.0, 2) if int(x) else None
            print(f"tower {tower} cta {cta}: setup {rel(r[1])} end {rel(r[127])}")
            print("  prod:", [rel(x) for x in r[2:34] if int(x)])
            print("  mma :", [rel(x) for x in r[48:56] if int(x)])
            print("  epi :", [rel(x) for x in r[64:72] if int(x)])
            print("  gbld:", [rel(x) for x in r[80:104] if int(x)])
            print("  cons:", [rel(x) for x in r[104:124] if int(x)])
    st = t[:, :, 0]
    t0all = int(st[st > 0].min())
    print("tower0 starts [%.2f, %.2f] ends [%.2f, %.2f]; tower1 starts [%.2f, %.2f] ends [%.2f, %.2f] us" % (
        (int(t[0, :, 0].min()) - t0all) / 1e3, (int(t[0, :, 0].max()) - t0all) / 1e3,
        (int(t[0, :, 127].min()) - t0all) / 1e3, (int(t[0, :, 127].max()) - t0all) / 1e3,
        (int(t[1, :, 0].min()) - t0all) / 1e3, (int(t[1, :, 0].max()) - t0all) / 1e3,
        (int(t[1, :, 127].min()) - t0all) / 1e3, (int(t[1, :, 127].max()) - t0all) / 1e3))
