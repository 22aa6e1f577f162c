import sys, torch, numpy as np, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import oracle
from ceo_firm_matching.scoring import score_topk, merge_topk
gen = torch.Generator().manual_seed(7)
_ = torch.randn(2048, 128, generator=gen); _ = torch.randn(2048, 128, generator=gen)
rows_all = F.normalize(torch.randn(2000, 60, generator=gen), dim=1)
cols_all = F.normalize(torch.randn(30001, 60, generator=gen), dim=1)
for (rlo, rhi) in [(0, 1000), (1000, 2000)]:
    r = rows_all[rlo:rhi].cuda(); c = cols_all.cuda()
    so, io = oracle.allpairs_topk(rows_all[rlo:rhi], cols_all, 100, 14.2857)
    s_full, i_full, fl = score_topk(r, c, 100, 14.2857, return_flags=True)
    print("full equal", bool((i_full.cpu() == io).all()), int(fl.sum()))
    parts = [score_topk(r, c[a:b], 100, 14.2857, col_offset=a) for a, b in ((0, 15001), (15001, 30001))]
    for (a, b), p in zip(((0, 15001), (15001, 30001)), parts):
        sp, ip = oracle.allpairs_topk(rows_all[rlo:rhi], cols_all[a:b], 100, 14.2857)
        print("  part", a, b, "equal", bool((p[1].cpu() == ip + a).all()))
    ms, mi = merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    bad = (mi.cpu() != io).any(1).nonzero().flatten()
    print("merged equal", bad.numel() == 0, "bad rows", bad[:5].tolist())
    if bad.numel():
        rr = int(bad[0]); print(mi[rr, :12].tolist()); print(io[rr, :12].tolist()); print(ms[rr,:6].tolist(), so[rr,:6].tolist())
