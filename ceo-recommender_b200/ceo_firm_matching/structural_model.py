"""Structural Distillation Network — drop-in for the reference's ``structural_model.py``.

Observables -> type logits per side -> frozen 5x5 BLM interaction matrix -> expected match
(reference ``ceo_firm_matching/structural_model.py:16-161``).  Same constructor, attributes
(``A`` buffer, ``firm_embeddings``, ``ceo_embeddings``, ``firm_tower``, ``ceo_tower``, ``config``) and
``state_dict`` keys; ``forward`` runs the fused tower ops and the fused softmax.A.softmax head.
"""
from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .structural_config import StructuralConfig

N_TYPES = 5


def _type_encoder(in_dim: int, latent: int, dropout: float) -> nn.Sequential:
    # structural_model.py:72-80: Linear-BN-ReLU-Dropout, Linear(latent,64)-ReLU, Linear(64,5)
    return nn.Sequential(
        nn.Linear(in_dim, latent), nn.BatchNorm1d(latent), nn.ReLU(), nn.Dropout(dropout),
        nn.Linear(latent, 64), nn.ReLU(), nn.Linear(64, N_TYPES))


class StructuralDistillationNet(nn.Module):
    def __init__(self, metadata: Dict, config: StructuralConfig):
        super().__init__()
        self.config = config
        # the structural constraint: a buffer, never updated by the optimiser (structural_model.py:50-53)
        self.register_buffer("A", torch.tensor(config.BLM_INTERACTION_MATRIX, dtype=torch.float32))
        e = config.EMBEDDING_DIM
        self.firm_embeddings = nn.ModuleList(nn.Embedding(n, e) for n in metadata["firm_cat_cards"])
        self.ceo_embeddings = nn.ModuleList(nn.Embedding(n, e) for n in metadata["ceo_cat_cards"])
        self.firm_tower = _type_encoder(metadata["n_firm_num"] + len(self.firm_embeddings) * e,
                                        config.LATENT_DIM, config.DROPOUT)
        self.ceo_tower = _type_encoder(metadata["n_ceo_num"] + len(self.ceo_embeddings) * e,
                                       config.LATENT_DIM, config.DROPOUT)
        self._handles = (
            ops.TowerHandle(self.firm_embeddings, self.firm_tower, (0, 4, 6), (1, None), (3, None), tower_id=0),
            ops.TowerHandle(self.ceo_embeddings, self.ceo_tower, (0, 4, 6), (1, None), (3, None), tower_id=1),
        )

    def logits(self, f_num, f_cat, c_num, c_cat) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(c_logits, f_logits)`` from the fused towers (structural_model.py:119-127)."""
        f_logits, c_logits = ops.run_towers(self._handles, [(f_num, f_cat), (c_num, c_cat)], self.training)
        return c_logits, f_logits

    def forward(self, f_num, f_cat, c_num, c_cat) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Returns ``(ceo_logits [B,5], firm_logits [B,5], expected_match [B,1])`` — CEO logits first."""
        c_logits, f_logits = self.logits(f_num, f_cat, c_num, c_cat)
        expected_match = ops.StructuralHeadFunction.apply(c_logits, f_logits, self.A)
        if not self.training:
            ops.raise_if_index_error(expected_match.device)
        return c_logits, f_logits, expected_match

    def distillation_loss(self, c_logits, f_logits, target_ceo, target_firm) -> torch.Tensor:
        """KL(batchmean) of both logit sets against the BLM posteriors, summed
        (structural_training.py:75-77) — loss and d/dlogits from one fused kernel."""
        return ops.StructuralKLFunction.apply(c_logits, f_logits, self.A, target_ceo, target_firm)

    def get_type_probabilities(self, f_num, f_cat, c_num, c_cat) -> Tuple[torch.Tensor, torch.Tensor]:
        """Softmax type probabilities ``(ceo_probs, firm_probs)``, each [B,5] (structural_model.py:145-161)."""
        c_logits, f_logits, _ = self.forward(f_num, f_cat, c_num, c_cat)
        return F.softmax(c_logits, dim=1), F.softmax(f_logits, dim=1)
