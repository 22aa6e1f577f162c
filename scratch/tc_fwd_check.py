"""Validate the tcgen05 forward stages against the mma.sync TF32 path and time fwd1..3 on config 4."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching import _native as N
from ceo_firm_matching.training import eager_step
dev = torch.device('cuda', 0)
lib = N.lib()
names = ["fwd1", "fwd2", "fwd3", "bwd1", "bwd2", "bwd3", "head", "emb", "reduce"]

def prof(model, batches, side):
    lib.cfm_profile_enable(1)
    with torch.cuda.stream(side):
        for i in range(len(batches)):
            eager_step(model, None, batches[i])
    torch.cuda.synchronize()
    ms = (C.c_double * 14)(); n = (C.c_int64 * 14)()
    N.check(lib.cfm_profile_read(ms, n, 14))
    lib.cfm_profile_enable(0)
    return {k: ms[i] / max(n[i], 1) for i, k in enumerate(names)}

side = torch.cuda.Stream()
for B in (65536, 1000, 77, 2):
    batches = bench.make_batches(3, B, dev, 1234)
    outs = {}
    for prec in ("tf32", "tf32_tc"):
        model = bench.build_model(dev, prec)
        with torch.cuda.stream(side):
            for i in range(2):
                eager_step(model, None, batches[i])
            loss, preds = model.forward_loss(*batches[2])
            f_raw, c_raw = model.encode_raw(*batches[2][:4]) if hasattr(model, "encode_raw") else (None, None)
        torch.cuda.synchronize()
        outs[prec] = (loss.detach().clone(), preds.detach().clone(),
                      {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k})
        if B == 65536:
            print(prec, {k: round(v, 4) for k, v in prof(model, batches, side).items()})
    a, b = outs["tf32"], outs["tf32_tc"]
    print(f"B={B}: loss {float(a[0]):.6f} vs {float(b[0]):.6f}  preds maxdiff {float((a[1]-b[1]).abs().max()):.3e}",
          " running maxdiff", max(float((a[2][k]-b[2][k]).abs().max()) for k in a[2]))
