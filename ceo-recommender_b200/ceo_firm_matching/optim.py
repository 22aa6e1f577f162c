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
        self._steps = {}                                # group index -> fp32 [n params] step counters

    def _state_for(self, group, params: List[torch.Tensor]) -> torch.Tensor:
        """Lazily create the state and return the group's step counters: ONE contiguous fp32 vector, of which every
        parameter's ``state["step"]`` is a distinct element (torch keeps a separate 0-dim tensor per parameter; after
        ``load_state_dict`` those copies are packed into one vector again)."""
        allp = group["params"]
        gi = next(i for i, g in enumerate(self.param_groups) if g is group)
        steps = self._steps.get(gi)
        index = {id(p): i for i, p in enumerate(allp)}
        packed = steps is not None and steps.numel() == len(allp) and all(
            "step" not in self.state.get(p, {}) or
            (self.state[p]["step"].dtype == torch.float32 and self.state[p]["step"].data_ptr() == steps.data_ptr() + 4 * i)
            for i, p in enumerate(allp))
        if not packed:
            dev = params[0].device
            new = torch.zeros(len(allp), dtype=torch.float32, device=dev)
            known = [float(self.state[p]["step"]) for p in allp if "step" in self.state.get(p, {})]
            if known:
                new.fill_(max(known))
            self._steps[gi] = steps = new
            for i, p in enumerate(allp):
                if "step" in self.state.get(p, {}):
                    self.state[p]["step"] = steps[i]
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                st["step"] = steps[index[id(p)]]
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return steps

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
            if group.get("weight_decay", 0) or group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("FusedAdam implements torch.optim.Adam with weight_decay=0, amsgrad=False, "
                                          "maximize=False (the reference's optimiser, training.py:32)")
            params = [p for p in group["params"] if p.grad is not None and p.numel() > 0]
            if not params:
                continue
            if len(params) != sum(1 for p in group["params"] if p.numel() > 0 and p.requires_grad) and self._steps:
                # one step counter drives the bias correction of the whole group (they advance together): a
                # parameter that skips steps would need torch's per-parameter count
                raise NotImplementedError("FusedAdam: every parameter of a group must receive a gradient at every step "
                                          "(per-parameter step counts are not tracked)")
            if any(not p.is_cuda for p in params):
                raise RuntimeError("FusedAdam runs on CUDA parameters only (this build has no CPU fallback)")
            steps = self._state_for(group, params)
            b1, b2 = group["betas"]
            with torch.cuda.device(params[0].device):
                N.check(N.lib().cfm_adam_step(self._records(params), len(params), N.ptr(steps), steps.numel(),
                                              float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                              self.variant, N.stream_ptr()))
        return loss
