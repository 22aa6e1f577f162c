import torch, struct
dev='cuda'
def hx(t): return struct.pack('>f', float(t)).hex()
for stepv in (1.0, 2.0, 6.0, 100.0):
    step = torch.tensor(stepv, device=dev)
    lr, b1, b2 = 1e-3, 0.9, 0.999
    bc1 = torch._foreach_pow(b1, [step])[0]; bc2 = torch._foreach_pow(b2, [step])[0]
    print("step", stepv, "foreach_pow", hx(bc1), hx(bc2), "| tensor.pow(float32 base)", hx(torch.tensor(b1, device=dev).pow(step)), hx(torch.tensor(b2, device=dev).pow(step)),
          "| double pow", hx(torch.tensor(b1 ** stepv)), hx(torch.tensor(b2 ** stepv)))
    x = [bc1.clone()]; torch._foreach_sub_(x, 1); a = x[0].clone()
    torch._foreach_div_(x, lr); b = x[0].clone()
    print("   sub", hx(a), "div lr", hx(b), "| true div", hx(a / torch.tensor(lr, device=dev)), "| mul recip(double)", hx(a * (1.0 / lr)), "| mul recip f32", hx(a * torch.tensor(1.0/lr, device=dev)))
    torch._foreach_reciprocal_(x); print("   recip", hx(x[0]), "| 1/b", hx(1.0 / b))
