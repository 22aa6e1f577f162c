"""Configuration of the Structural Distillation Network — drop-in for the reference's
``structural_config.py:12-82`` (same field names, defaults and the frozen BLM interaction matrix)."""
from dataclasses import dataclass, field
from typing import List

import torch

from .config import pick_device

# 5x5 interaction matrix A (rows = CEO types, columns = firm classes), structural_config.py:41-47
_BLM_A = (
    (-0.5, -0.3, 0.0, 0.1, 0.2),
    (-0.2, -0.1, 0.1, 0.3, 0.4),
    (0.0, 0.2, 0.4, 0.6, 0.7),
    (0.1, 0.4, 0.7, 0.9, 1.1),
    (0.3, 0.6, 0.9, 1.2, 1.5),
)


def _prob_cols(side: str) -> List[str]:
    return [f"prob_{side}_{k}" for k in range(1, 6)]


@dataclass
class StructuralConfig:
    DEVICE: torch.device = field(default_factory=pick_device)
    DATA_PATH: str = "Data/blm_posteriors.csv"
    OUTPUT_PATH: str = "./Output/Structural_Distillation"

    EPOCHS: int = 50
    LEARNING_RATE: float = 0.001
    BATCH_SIZE: int = 256
    DROPOUT: float = 0.2
    LATENT_DIM: int = 128
    EMBEDDING_DIM: int = 8

    BLM_INTERACTION_MATRIX: List[List[float]] = field(default_factory=lambda: [list(r) for r in _BLM_A])

    CEO_PROB_COLS: List[str] = field(default_factory=lambda: _prob_cols("ceo"))
    FIRM_PROB_COLS: List[str] = field(default_factory=lambda: _prob_cols("firm"))

    CEO_NUMERIC_COLS: List[str] = field(default_factory=lambda: ["Age", "tenure"])
    CEO_CAT_COLS: List[str] = field(default_factory=lambda: [
        "Gender", "maxedu", "ivy", "m", "Output", "Throghput", "Peripheral"])
    FIRM_NUMERIC_COLS: List[str] = field(default_factory=lambda: [
        "ind_firms_60w", "non_competition_score", "boardindpw", "boardsizew", "busyw", "pct_blockw",
        "logatw", "exp_roa", "rdintw", "capintw", "leverage", "divyieldw"])
    FIRM_CAT_COLS: List[str] = field(default_factory=lambda: ["compindustry", "ba_state", "rd_control", "dpayer"])

    @property
    def all_cols(self) -> List[str]:
        """Raw columns needed from the CSV ('tenure' is derived from fiscalyear - ceo_year)."""
        cols = (self.CEO_CAT_COLS + self.FIRM_NUMERIC_COLS + self.FIRM_CAT_COLS + self.CEO_PROB_COLS
                + self.FIRM_PROB_COLS + ["Age", "fiscalyear", "ceo_year"])
        return list(set(cols))
