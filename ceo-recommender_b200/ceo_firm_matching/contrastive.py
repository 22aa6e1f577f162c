"""Contrastive two-tower extension — drop-in for the hot-path part of the reference's ``contrastive.py``:
``ContrastiveCEOFirmMatcher`` (:21-99), ``info_nce_loss`` (:102-138), ``train_contrastive`` (:197-272, InfoNCE
branch) and ``compute_retrieval_metrics`` (:275-332).  The in-batch similarity matrix and its softmax-cross-entropy
backward run as TMA-fed tcgen05/TMEM GEMMs with the exponentials fused into the epilogue; the ``[B,B]`` matrix is
never written to memory, so the batch is bounded by time, not by a 17 GB tensor.  Semi-hard triplet mining
(:141-194, a Python per-row loop) is out of scope.
"""
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader

from . import ops
from .batching import device_batches
from .config import Config
from .model import CEOFirmMatcher
from .training import BATCH_KEYS


def _projector(dim: int) -> nn.Sequential:
    # contrastive.py:41-50: Linear(D, D) - ReLU - Linear(D, D // 2)
    return nn.Sequential(nn.Linear(dim, dim), nn.ReLU(), nn.Linear(dim, dim // 2))


class ContrastiveCEOFirmMatcher(nn.Module):
    """Base two-tower model plus one projection head per side for the contrastive objective."""

    def __init__(self, metadata: Dict[str, int], config: Config):
        super().__init__()
        self.base_model = CEOFirmMatcher(metadata, config)
        self.config = config
        self.firm_projector = _projector(config.LATENT_DIM)
        self.ceo_projector = _projector(config.LATENT_DIM)

    def _unit_latents(self, f_numeric, f_cat, c_numeric, c_cat):
        u, v = self.base_model.encode_raw(f_numeric, f_cat, c_numeric, c_cat)
        # F.normalize semantics (eps 1e-12) at this call site, unlike the base model (contrastive.py:64,70)
        return ops.CosineHeadFunction.apply(u, v, self.base_model.logit_scale, 1e-12, True)

    def get_embeddings(self, f_numeric, f_cat, c_numeric, c_cat) -> Tuple[torch.Tensor, torch.Tensor]:
        """L2-normalised latents of both towers (contrastive.py:52-72)."""
        _, u_hat, v_hat = self._unit_latents(f_numeric, f_cat, c_numeric, c_cat)
        return u_hat, v_hat

    def forward(self, f_numeric, f_cat, c_numeric, c_cat):
        """``(match_score [B,1], firm_proj [B,D//2], ceo_proj [B,D//2])`` (contrastive.py:74-99)."""
        score, u_hat, v_hat = self._unit_latents(f_numeric, f_cat, c_numeric, c_cat)
        # Both projection heads (Linear - ReLU - Linear) and the F.normalize that follows them run as ONE launch of
        # the library's projector kernel; the nn.Sequential sub-modules stay in place for state_dict / direct callers
        # and only hand over their parameters, like the towers.
        firm_proj, ceo_proj = ops.projector_heads([(u_hat, self.firm_projector), (v_hat, self.ceo_projector)], 1e-12)
        return score, firm_proj, ceo_proj


def info_nce_loss(firm_proj: torch.Tensor, ceo_proj: torch.Tensor, temperature: float = 0.07) -> torch.Tensor:
    """Symmetric InfoNCE over in-batch negatives (contrastive.py:102-138): rows of ``firm_proj`` / ``ceo_proj`` are
    unit vectors, positives sit on the diagonal of ``S = F C^T / T``; returns 0 for ``B <= 1``.
    Computed in bf16 on the tensor cores with fp32 accumulation (tolerance: 1e-3 relative on the loss).
    Inputs must be unit rows and ``temperature >= 0.012`` (both checked): the kernels subtract the fixed maximum 1/T."""
    B = firm_proj.size(0)
    if B <= 1:
        return torch.tensor(0.0, device=firm_proj.device)
    # The kernels use the fixed softmax shift 1/T (|s| <= 1 for unit rows), which makes partial sums addable across
    # column chunks and GPUs; rows that are not unit vectors would overflow it, and a temperature below ~0.012 lets
    # exp((s - 1)/T) underflow fp32 for s < 0.  Checked here (one small reduction + host read) unless a CUDA graph is
    # being captured.
    if not 0.012 <= float(temperature):
        raise ValueError("info_nce_loss: temperature must be >= 0.012 for the fixed-shift softmax (reference default 0.07)")
    if not torch.cuda.is_current_stream_capturing():
        nmax = float(torch.maximum(firm_proj.detach().norm(dim=1).max(), ceo_proj.detach().norm(dim=1).max()))
        if not nmax <= 1.0 + 1e-3:
            raise ValueError(f"info_nce_loss expects L2-normalised rows (contrastive.py:96-97); largest row norm is {nmax:.4f}")
    return ops.InfoNCEFunction.apply(firm_proj, ceo_proj, float(temperature))


def train_contrastive(train_loader: DataLoader, val_loader: DataLoader, metadata: Dict[str, int], config: Config,
                      contrastive_weight: float = 0.3, temperature: float = 0.07,
                      use_triplet: bool = False) -> ContrastiveCEOFirmMatcher:
    """Adam on ``(1 - a) * weighted MSE + a * InfoNCE`` (contrastive.py:197-272)."""
    if use_triplet:
        raise NotImplementedError("semi-hard triplet mining (contrastive.py:141-194) is out of scope of this build")
    device = torch.device(config.DEVICE)
    model = ContrastiveCEOFirmMatcher(metadata, config).to(device)
    from .optim import FusedAdam
    optimizer = FusedAdam(model.parameters(), lr=config.LEARNING_RATE)   # optim.Adam's state/arithmetic, one launch
    print(f"Training Contrastive Two-Tower on {config.DEVICE}")
    print(f"  Contrastive weight: {contrastive_weight}")
    print(f"  Temperature: {temperature}")
    print("  Loss type: InfoNCE")
    for epoch in range(config.EPOCHS):
        model.train()
        totals = torch.zeros(3, device=device)
        n_batches = 0
        for f_num, f_cat, c_num, c_cat, target, weights in device_batches(train_loader, BATCH_KEYS, device):
            optimizer.zero_grad(set_to_none=True)
            score, firm_proj, ceo_proj = model(f_num, f_cat, c_num, c_cat)
            mse_loss = (weights * (score - target) ** 2).mean()
            cl_loss = info_nce_loss(firm_proj, ceo_proj, temperature)
            loss = (1 - contrastive_weight) * mse_loss + contrastive_weight * cl_loss
            loss.backward()
            optimizer.step()
            totals += torch.stack([loss.detach(), mse_loss.detach(), cl_loss.detach()])
            n_batches += 1
        ops.raise_if_index_error(device)
        if epoch % 5 == 0:
            avg = (totals / max(n_batches, 1)).tolist()
            print(f"  Epoch {epoch}: Loss={avg[0]:.4f} (MSE={avg[1]:.4f}, CL={avg[2]:.4f})")
    return model


def compute_retrieval_metrics(model: ContrastiveCEOFirmMatcher, data_dict: Dict, config: Config,
                              top_k: int = 10) -> Dict[str, float]:
    """recall@1/5/k, MRR and median rank of the true CEO among all CEOs for every firm (contrastive.py:275-332).
    The rank of the diagonal is counted directly (``1 + #{j : s_ij > s_ii}``) instead of sorting every row, and
    the reference's 5000-row memory cap is kept only as the default window."""
    from .scoring import diagonal_ranks
    device = torch.device(config.DEVICE)
    model.eval()
    with torch.no_grad():
        ins = [data_dict[k].to(device) for k in ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat")]
        firm_emb, ceo_emb = model.get_embeddings(*ins)
        n = min(firm_emb.size(0), 5000)
        ranks = diagonal_ranks(firm_emb[:n].contiguous(), ceo_emb[:n].contiguous()).cpu().numpy()
    return {
        "recall@1": float(np.mean(ranks <= 1)), "recall@5": float(np.mean(ranks <= 5)),
        "recall@10": float(np.mean(ranks <= top_k)), "MRR": float(np.mean(1.0 / ranks)),
        "median_rank": float(np.median(ranks)),
    }
