"""Fused dense Adam for the two-tower models: `torch.optim.Adam(capturable=True)` semantics —
same state (`step`, `exp_avg`, `exp_avg_sq`), same arithmetic operation by operation — in ONE kernel launch over all
parameter tensors (`cfm_adam_step`) instead of torch's ~12 multi-tensor passes.  With 1M-row embedding tables the
optimiser is the largest mover of bytes in a training step (training.py:32,55); this pass moves 28 B per parameter.

No fallback: parameters and gradients must be dense, contiguous fp32 CUDA tensors."""
from typing import List, Tuple

import torch

from . import _native as N


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 variant: int = 0):
        if lr <= 0 or not 0.5 < betas[0] < 1 or not 0 <= betas[1] < 1 or eps < 0:
            raise ValueError("FusedAdam needs lr > 0, 0.5 < beta1 < 1, 0 <= beta2 < 1, eps >= 0")
        # the keys torch.optim.Adam keeps, so param_groups / state_dict stay interchangeable
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                        capturable=True, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self.variant = variant

    def _state_for(self, group, params: List[torch.Tensor]):
        step = None
        for p in group["params"]:                      # one device step counter per group, shared by its parameters
            st = self.state.get(p)
            if st and "step" in st:
                step = st["step"]
                break
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                if step is None:
                    step = torch.zeros((), dtype=torch.float32, device=p.device)
                st["step"] = step
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            elif st["step"] is not step:                # after load_state_dict: re-alias the per-parameter copies
                if st["step"].device != p.device or st["step"].dtype != torch.float32:
                    st["step"] = st["step"].to(device=p.device, dtype=torch.float32)
                if step is None:
                    step = st["step"]
                st["step"] = step
        return step

    def _records(self, params: List[torch.Tensor]):
        arr = (N.AdamTensor * len(params))()
        for i, p in enumerate(params):
            st = self.state[p]
            g = p.grad
            for t, what in ((p, "parameter"), (g, "gradient"), (st["exp_avg"], "exp_avg"), (st["exp_avg_sq"], "exp_avg_sq")):
                if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and not t.is_sparse):
                    raise RuntimeError(f"FusedAdam: {what} must be a dense contiguous fp32 CUDA tensor (no fallback)")
            arr[i].param, arr[i].grad = p.data_ptr(), g.data_ptr()
            arr[i].exp_avg, arr[i].exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            arr[i].numel = p.numel()
        return arr

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None and p.numel() > 0]
            if not params:
                continue
            step = self._state_for(group, params)
            b1, b2 = group["betas"]
            with torch.cuda.device(params[0].device):
                N.check(N.lib().cfm_adam_step(self._records(params), len(params), N.ptr(step), float(group["lr"]),
                                              float(b1), float(b2), float(group["eps"]), self.variant,
                                              N.stream_ptr()))
        return loss
