"""Hot-path part of the reference's ``analytical_extensions.py``: ``generate_counterfactuals`` (:405-523), the
"what if CEO X ran firm Y" cross-match.  The reference encodes every row on the CPU, forms the full
``[firms, CEOs]`` score matrix with ``torch.mm`` (:471) and argsorts every row in a Python loop (:483), capped at
200 firms x 1000 CEOs for memory.  Here the embeddings come from the fused eval forward and the ranking questions the
table asks (best CEO, worst CEO, rank of the actual CEO) go to the tcgen05 all-pairs kernels: ``scoring.score_topk``
(best / worst, no score matrix in memory) and ``scoring.target_ranks`` (rank counted, no sort), so the caps are only
defaults.  The rest of the module (plots, econometrics) is out of scope of this build.
"""
from typing import Dict, Optional

import numpy as np
import pandas as pd
import torch
import torch.nn as nn

from . import ops
from .scoring import score_topk, target_ranks


def _unit_latents(model: nn.Module, data_dict: Dict[str, torch.Tensor], device):
    """L2-normalised tower outputs (F.normalize semantics, analytical_extensions.py:439-451) from the fused towers."""
    ins = [data_dict[k].to(device) for k in ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat")]
    if hasattr(model, "encode"):
        return model.encode(*ins)
    base = model.base_model if hasattr(model, "base_model") else model
    u, v = base.encode_raw(*ins)
    _, u_hat, v_hat = ops.CosineHeadFunction.apply(u, v, base.logit_scale, 1e-12, True)
    return u_hat, v_hat


def generate_counterfactuals(model: nn.Module, data_dict: Dict[str, torch.Tensor], df: pd.DataFrame,
                             firm_id_col: str = "gvkey", ceo_id_col: str = "match_exec_id", top_k: int = 10,
                             device: Optional[str] = None, max_firms: Optional[int] = 200,
                             max_ceos: Optional[int] = 1000) -> pd.DataFrame:
    """Counterfactual match table, one row per firm: actual CEO and its rank among all candidate CEOs, best and worst
    CEO with their scores (same columns and prints as the reference).  ``max_firms`` / ``max_ceos`` default to the
    reference's caps; ``None`` lifts them.  Ties are ordered (score desc, index asc) - the reference's
    ``np.argsort`` leaves them unspecified."""
    device = torch.device(device) if device is not None else next(model.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("ceo_firm_matching (B200 build) scores on CUDA only; pass device='cuda'")
    model.eval()
    with torch.no_grad():
        u_firm, v_ceo = _unit_latents(model, data_dict, device)
        base = model.base_model if hasattr(model, "base_model") else model
        logit_scale = float(base.logit_scale.exp()) if hasattr(base, "logit_scale") else 14.3

    df = df.reset_index(drop=True)
    latest = df.groupby(firm_id_col)["fiscalyear"].idxmax() if "fiscalyear" in df.columns else df.index
    firm_indices = (df.loc[latest].index.values if hasattr(latest, "values") else np.asarray(latest))[:max_firms]
    ceo_unique = df.drop_duplicates(subset=ceo_id_col, keep="last") if ceo_id_col in df.columns else df
    ceo_indices = ceo_unique.index.values[:max_ceos]

    firm_embs = u_firm[torch.as_tensor(firm_indices, device=device)].contiguous()      # [F, D]
    ceo_embs = v_ceo[torch.as_tensor(ceo_indices, device=device)].contiguous()         # [C, D]
    print("\n=== Counterfactual Analysis ===")
    print(f"Cross-matching {len(firm_indices)} firms × {len(ceo_indices)} CEOs")

    with torch.no_grad():
        best_s, best_i = score_topk(firm_embs, ceo_embs, 1, logit_scale)               # argsort(-scores)[0]
        worst_s, worst_i = score_topk(-firm_embs, ceo_embs, 1, logit_scale)            # argsort(-scores)[-1]
        has_ids = ceo_id_col in df.columns
        actual_pos = np.full(len(firm_indices), -1, dtype=np.int64)
        if has_ids:
            first_pos = {}
            for pos, cid in enumerate(df.loc[ceo_indices, ceo_id_col].values):       # first occurrence, as .argmax()
                first_pos.setdefault(cid, pos)
            actual_pos = np.array([first_pos.get(c, -1) for c in df.loc[firm_indices, ceo_id_col].values], dtype=np.int64)
        ranks = np.zeros(len(firm_indices), dtype=np.int64)
        known = np.nonzero(actual_pos >= 0)[0]
        if known.size:
            kk = torch.as_tensor(known, device=device)
            ranks[known] = target_ranks(firm_embs[kk].contiguous(), ceo_embs,
                                        torch.as_tensor(actual_pos[known], device=device)).cpu().numpy()
    best_s, best_i = best_s[:, 0].cpu().numpy(), best_i[:, 0].cpu().numpy()
    worst_s, worst_i = -worst_s[:, 0].cpu().numpy(), worst_i[:, 0].cpu().numpy()

    results = []
    for i, firm_idx in enumerate(firm_indices):
        actual_ceo = df.loc[firm_idx, ceo_id_col] if has_ids else None
        actual_match = df.loc[firm_idx, "match_means"] if "match_means" in df.columns else None
        results.append({
            "firm_id": df.loc[firm_idx, firm_id_col] if firm_id_col in df.columns else firm_idx,
            "actual_ceo": actual_ceo,
            "actual_match": actual_match,
            "actual_rank": int(ranks[i]) if actual_pos[i] >= 0 else None,
            "best_ceo": df.loc[ceo_indices[best_i[i]], ceo_id_col] if has_ids else int(best_i[i]),
            "best_score": float(best_s[i]),
            "worst_ceo": df.loc[ceo_indices[worst_i[i]], ceo_id_col] if has_ids else int(worst_i[i]),
            "worst_score": float(worst_s[i]),
            "score_range": float(best_s[i] - worst_s[i]),
            "match_improvement": float(best_s[i]) - (actual_match if actual_match else 0),
        })
    cf_df = pd.DataFrame(results)
    print("\nMatch Improvement Statistics:")
    print(f"  Mean actual rank: {cf_df['actual_rank'].mean():.1f} / {len(ceo_indices)}")
    print(f"  Median improvement: {cf_df['match_improvement'].median():.3f}")
    print(f"  % firms with better CEO available: {(cf_df['match_improvement'] > 0).mean():.1%}")
    return cf_df
