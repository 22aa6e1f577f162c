"""Two-tower CEO-firm matcher — drop-in for the reference's ``model.py``.

Same constructor, attribute names and ``state_dict`` layout as ``CEOFirmMatcher``
(reference ``ceo_firm_matching/model.py:14-89``): ``firm_embeddings``, ``ceo_embeddings``
(``nn.ModuleList`` of ``nn.Embedding``), ``firm_tower``, ``ceo_tower`` (``nn.Sequential``)
and the scalar ``logit_scale``.  The sub-modules are real torch modules so the scripts that
reach into them keep working; ``forward`` does not call them — it hands their parameters to the
fused CUDA ops (gather+concat+MLP towers, then normalise/cosine/scale).
"""
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .config import Config


def _mlp(in_dim: int, out_dim: int) -> nn.Sequential:
    # model.py:37-47: 64 -> 32 hidden units, BatchNorm + ReLU + Dropout(0.1) after each hidden Linear
    layers = []
    width = in_dim
    for hidden in (64, 32):
        layers += [nn.Linear(width, hidden), nn.BatchNorm1d(hidden), nn.ReLU(), nn.Dropout(0.1)]
        width = hidden
    layers.append(nn.Linear(width, out_dim))
    return nn.Sequential(*layers)


class CEOFirmMatcher(nn.Module):
    """Encodes firm and CEO features with separate towers and scores the pair by scaled cosine."""

    def __init__(self, metadata: Dict[str, int], config: Config):
        super().__init__()
        e_firm, e_ceo = config.EMBEDDING_DIM_LARGE, config.EMBEDDING_DIM_MEDIUM
        self.firm_embeddings = nn.ModuleList(nn.Embedding(n, e_firm) for n in metadata["firm_cat_counts"])
        self.ceo_embeddings = nn.ModuleList(nn.Embedding(n, e_ceo) for n in metadata["ceo_cat_counts"])
        self.firm_tower = _mlp(metadata["n_firm_numeric"] + len(self.firm_embeddings) * e_firm, config.LATENT_DIM)
        self.ceo_tower = _mlp(metadata["n_ceo_numeric"] + len(self.ceo_embeddings) * e_ceo, config.LATENT_DIM)
        # learnable temperature of the cosine score, initialised to ln(1/0.07) (model.py:65)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self._handles = (
            ops.TowerHandle(self.firm_embeddings, self.firm_tower, (0, 4, 8), (1, 5), (3, 7), tower_id=0),
            ops.TowerHandle(self.ceo_embeddings, self.ceo_tower, (0, 4, 8), (1, 5), (3, 7), tower_id=1),
        )

    # ---- fused building blocks ----------------------------------------------------------
    def encode_raw(self, f_numeric, f_cat, c_numeric, c_cat):
        """Un-normalised latents ``(u_firm, v_ceo)`` of both towers (model.py:69-76)."""
        return ops.run_towers(self._handles, [(f_numeric, f_cat), (c_numeric, c_cat)], self.training)

    def forward(self, f_numeric, f_cat, c_numeric, c_cat):
        u, v = self.encode_raw(f_numeric, f_cat, c_numeric, c_cat)
        # model.py:79-87: plain division by the norm (no eps), dot, times exp(logit_scale)
        score = ops.CosineHeadFunction.apply(u, v, self.logit_scale, 0.0, False)
        if not self.training:
            ops.raise_if_index_error(score.device)
        return score

    def forward_loss(self, f_numeric, f_cat, c_numeric, c_cat, target, weights):
        """Fused training objective of training.py:46-52: returns ``(loss, preds)`` with
        ``loss = (weights * (preds - target)**2).mean()`` computed inside the head kernel."""
        u, v = self.encode_raw(f_numeric, f_cat, c_numeric, c_cat)
        return ops.CosineMSEFunction.apply(u, v, self.logit_scale, target, weights, 0.0)

    def set_precision(self, precision: str = "fp32") -> None:
        """``"fp32"``: tower products are 3xTF32 error-compensated (fp32-class, the parity default);
        ``"tf32"``: one TF32 tensor-core pass (10-bit mantissa operands, fp32 accumulation; ~1e-3 relative)."""
        code = {"fp32": 0, "tf32": 1}[precision]
        for h in self._handles:
            h.precision = code

    # ---- dense-gradient bookkeeping for large embedding tables --------------------------
    def use_persistent_table_grads(self, enable: bool = True) -> None:
        """Keep every embedding table's dense ``.grad`` allocated and re-zero only the rows touched by the
        previous step (exactly the dense gradient torch would produce, without a full-table memset)."""
        for h in self._handles:
            h.table_grads = ops.PersistentTableGrads(h) if enable else None
        if enable:
            ops.JointTableGrads(self._handles)   # both towers' gradients go through one radix sort per step

    def rezero_table_grads(self) -> None:
        """Zero the table-gradient rows written by the last backward (call after ``optimizer.step()``)."""
        for h in self._handles:
            if h.table_grads is not None:
                h.table_grads.rezero()

    def zero_grad_fast(self) -> None:
        """``optimizer.zero_grad()`` for the fused loop: dense parameters drop their grads, tables re-zero sparsely."""
        tables = set()
        for h in self._handles:
            if h.table_grads is not None:
                h.table_grads.rezero()
                tables.update(id(e.weight) for e in h.embeddings)
        for p in self.parameters():
            if id(p) not in tables:
                p.grad = None
