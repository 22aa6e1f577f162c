"""Reusable tower encoder for every model that keeps the reference's tower layout, and batched grid scoring.

The reference's extension models repeat the same ``encode`` body — gather the categorical embeddings, concatenate
with the numerics, run the 3-layer tower, L2-normalise (``multitask_model.py:127-147``, ``industry_model.py:104-127``,
``contrastive.py:52-77``) — and ``visualization.py:99-123`` scores a 2-D grid of counterfactual inputs with one
single-row forward per cell.  Both map onto the fused tower kernels:

* ``encode(module, f_numeric, f_cat, c_numeric, c_cat)`` works on ANY ``nn.Module`` exposing ``firm_embeddings`` /
  ``ceo_embeddings`` (``nn.ModuleList`` of ``nn.Embedding``) and ``firm_tower`` / ``ceo_tower`` (``nn.Sequential`` of
  three ``Linear`` layers with optional ``BatchNorm1d`` / ``Dropout`` after the first two) — the product classes and the
  reference's own extension models alike; differentiable, train/eval aware;
* ``score_grid`` builds all cells of an interaction heat-map as ONE batch and runs one forward.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops


def _tower_layout(seq: nn.Sequential) -> Tuple[Tuple[int, int, int], Tuple[Optional[int], Optional[int]],
                                               Tuple[Optional[int], Optional[int]]]:
    """Indices of the three Linear layers and of the BatchNorm1d / Dropout modules that follow the first two."""
    lin = [i for i, m in enumerate(seq) if isinstance(m, nn.Linear)]
    if len(lin) != 3:
        raise ValueError(f"encode(): expected a tower with three Linear layers, found {len(lin)}")
    bn: List[Optional[int]] = []
    drop: List[Optional[int]] = []
    for a, b in ((lin[0], lin[1]), (lin[1], lin[2])):
        between = list(range(a + 1, b))
        bns = [i for i in between if isinstance(seq[i], nn.BatchNorm1d)]
        drops = [i for i in between if isinstance(seq[i], nn.Dropout)]
        relus = [i for i in between if isinstance(seq[i], nn.ReLU)]
        other = [i for i in between if i not in bns + drops + relus]
        if len(relus) != 1 or len(bns) > 1 or len(drops) > 1 or other:
            raise ValueError("encode(): tower blocks must be Linear -> [BatchNorm1d] -> ReLU -> [Dropout]")
        bn.append(bns[0] if bns else None)
        drop.append(drops[0] if drops else None)
    if bn[0] is None:
        raise ValueError("encode(): the first tower block needs its BatchNorm1d (reference layout)")
    if lin[2] != len(seq) - 1:
        raise ValueError("encode(): the tower must end with its third Linear layer")
    return tuple(lin), tuple(bn), tuple(drop)


def tower_handles(module: nn.Module) -> Tuple["ops.TowerHandle", "ops.TowerHandle"]:
    """The (firm, CEO) ``TowerHandle`` pair of ``module`` (built once, cached on the module)."""
    cached = getattr(module, "_handles", None) or getattr(module, "_cfm_handles", None)
    if cached is not None:
        return cached
    hs = []
    for tid, (emb, tower) in enumerate((("firm_embeddings", "firm_tower"), ("ceo_embeddings", "ceo_tower"))):
        seq = getattr(module, tower)
        lin, bn, drop = _tower_layout(seq)
        hs.append(ops.TowerHandle(getattr(module, emb), seq, lin, bn, drop, tower_id=tid))
    object.__setattr__(module, "_cfm_handles", tuple(hs))
    return module._cfm_handles


def encode(module: nn.Module, f_numeric: torch.Tensor, f_cat: torch.Tensor, c_numeric: torch.Tensor,
           c_cat: torch.Tensor, normalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(u_firm, v_ceo)`` of ``module``'s towers on the fused CUDA path; ``normalize`` applies ``F.normalize(dim=1)``
    as the reference's ``encode`` methods do (``multitask_model.py:139,145``)."""
    u, v = ops.run_towers(tower_handles(module), [(f_numeric, f_cat), (c_numeric, c_cat)], module.training)
    if normalize:
        # F.normalize semantics (x / max(||x||, 1e-12)) from the cosine-head kernel's unit-latent outputs
        zero = torch.zeros((), device=u.device)
        if u.shape == v.shape:
            _, u, v = ops.CosineHeadFunction.apply(u, v, zero, 1e-12, True)
        else:                                   # towers of different widths: one launch per side
            _, u, _ = ops.CosineHeadFunction.apply(u, u, zero, 1e-12, True)
            _, v, _ = ops.CosineHeadFunction.apply(v, v, zero, 1e-12, True)
    return u, v


def score_grid(model: nn.Module, base: Sequence[torch.Tensor], x_spec: Tuple[str, int], x_vals: Sequence[float],
               y_spec: Tuple[str, int], y_vals: Sequence[float]) -> torch.Tensor:
    """Interaction heat-map of ``visualization.py:99-123``: ``heat[i, j] = model(inputs with y := y_vals[i], x := x_vals[j])``.

    ``base`` = ``(f_numeric [1,nf], f_cat [1,kf], c_numeric [1,nc], c_cat [1,kc])`` (the reference uses the column means
    and modes); a spec is ``(feature type, column)`` with the type one of ``firm_numeric``, ``ceo_numeric``,
    ``firm_cat``, ``ceo_cat``.  All ``len(y_vals) * len(x_vals)`` cells go through the towers as one batch in eval
    mode (BatchNorm running statistics: rows do not interact, so the cells equal the one-row forwards of the reference).
    """
    slot = {"firm_numeric": 0, "firm_cat": 1, "ceo_numeric": 2, "ceo_cat": 3}
    ny, nx = len(y_vals), len(x_vals)
    cells = [t.detach().reshape(1, -1).expand(ny * nx, -1).clone() for t in base]
    dev = cells[0].device
    xs = torch.as_tensor(list(x_vals), dtype=torch.float64, device=dev).repeat(ny)               # x varies fastest
    ys = torch.as_tensor(list(y_vals), dtype=torch.float64, device=dev).repeat_interleave(nx)
    for (ftype, col), vals in ((x_spec, xs), (y_spec, ys)):                                         # y written last, as
        tgt = cells[slot[ftype]]                                                                    # in the reference
        tgt[:, col] = vals.to(tgt.dtype)
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            score = model(*cells)
    finally:
        model.train(was_training)
    return score.reshape(ny, nx)
