import sys, torch
sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching import _native as N
dev = torch.device('cuda', 0)
out = torch.zeros(2, dtype=torch.int64, device=dev)
lib = N.lib()
for (M, Nn) in [(128, 64), (64, 32), (128, 32), (128, 256), (64, 208), (128, 128), (64, 8)]:
    for nacc in (1, 2, 4):
        if nacc * Nn > 512: continue
        res = []
        for reps in (8, 64, 256):
            N.check(lib.cfm_tc_mma_probe(N.ptr(out), M, Nn, reps, nacc, N.stream_ptr()))
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            res.append((reps, o[0], o[1]))
        per = (res[2][1] - res[1][1]) / (256 - 64)
        print(f"M{M} N{Nn} nacc{nacc}: {res}  -> {per:.1f} cycles/MMA steady, issue {(res[2][2]-res[1][2])/192:.1f}")
