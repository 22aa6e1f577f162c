import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from ceo_firm_matching import ops
import oracle
def unit(R, D, seed):
    g = torch.Generator().manual_seed(seed); return F.normalize(torch.randn(R, D, generator=g), dim=1)
for B, D in [(2, 16), (3, 16), (2, 64), (5, 30)]:
    f0, c0 = unit(B, D, 10 + B), unit(B, D, 20 + B)
    c0 = F.normalize(0.6 * f0 + 0.8 * c0, dim=1)
    fb, cb = ops.pack_bf16(f0.cuda()), ops.pack_bf16(c0.cuda())
    rs_f, diag = ops.infonce_rowsum(fb, cb, 0.07)
    rs_c, _ = ops.infonce_rowsum(cb, fb, 0.07, want_diag=False)
    s = fb.float().cpu().double() @ cb.float().cpu().double().t()
    E = torch.exp((s - 1) / 0.07)
    print(B, D, 'rs_f', rs_f.cpu().tolist(), E.sum(1).tolist())
    print('   rs_c', rs_c.cpu().tolist(), E.sum(0).tolist())
    print('   diag', diag.cpu().tolist(), s.diag().tolist())
    fq, cq = f0.bfloat16().float(), c0.bfloat16().float()
    print('   oracle', float(oracle.info_nce(fq, cq, 0.07)), 'formula', float((torch.log(E.sum(1)) + torch.log(E.sum(0)) + 2/0.07 - 2*s.diag()/0.07).sum() / (2*B)))
