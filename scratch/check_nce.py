import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import oracle
from ceo_firm_matching.contrastive import info_nce_loss
dev = torch.device('cuda', 0)
for B, D in [(200, 30), (200, 64), (200, 128), (129, 60), (300, 16), (127, 30), (640, 60)]:
    gen = torch.Generator().manual_seed(2)
    fp = F.normalize(torch.randn(B, D, generator=gen), dim=1)
    cp = F.normalize(0.6 * fp + 0.8 * torch.randn(B, D, generator=gen), dim=1)
    fd, cd = fp.to(dev).requires_grad_(True), cp.to(dev).requires_grad_(True)
    cl = info_nce_loss(fd, cd, 0.07); cl.backward()
    fo, co = fp.clone().requires_grad_(True), cp.clone().requires_grad_(True)
    clo = oracle.info_nce(fo, co, 0.07); clo.backward()
    ef = float((fd.grad.cpu() - fo.grad).abs().max() / fo.grad.abs().max())
    ec = float((cd.grad.cpu() - co.grad).abs().max() / co.grad.abs().max())
    print(B, D, "loss rel", abs(float(cl) - float(clo)) / float(clo), "d_firm", ef, "d_ceo", ec)
from ceo_firm_matching import _native as N
for mask in (0, 1):
    N.lib().cfm_simtile_set_poly(mask)
    B, D = 200, 30
    gen = torch.Generator().manual_seed(2)
    fp = F.normalize(torch.randn(B, D, generator=gen), dim=1)
    cp = F.normalize(0.6 * fp + 0.8 * torch.randn(B, D, generator=gen), dim=1)
    fd, cd = fp.to(dev).requires_grad_(True), cp.to(dev).requires_grad_(True)
    cl = info_nce_loss(fd, cd, 0.07); cl.backward()
    fo, co = fp.clone().requires_grad_(True), cp.clone().requires_grad_(True)
    clo = oracle.info_nce(fo, co, 0.07); clo.backward()
    # reference with bf16-rounded operands (what the kernel is given): isolates kernel error from operand rounding
    fb, cb = fp.bfloat16().float().requires_grad_(True), cp.bfloat16().float().requires_grad_(True)
    clb = oracle.info_nce(fb, cb, 0.07); clb.backward()
    ef = float((fd.grad.cpu() - fo.grad).abs().max() / fo.grad.abs().max())
    eb = float((fd.grad.cpu() - fb.grad).abs().max() / fb.grad.abs().max())
    print("poly", mask, "d_firm vs fp32 oracle", ef, "vs oracle on bf16-rounded operands", eb)
