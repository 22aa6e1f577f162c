"""Configuration of the Structural Distillation Network — drop-in for the reference's
``structural_config.py:12-82`` (same field names, defaults and the frozen BLM interaction matrix; the column
vocabulary comes from ``schema.py``)."""
from dataclasses import field, make_dataclass
from typing import List

import torch

from . import schema
from .config import pick_device

# 5x5 interaction matrix A (rows = CEO types, columns = firm classes), structural_config.py:41-47
_BLM_A = (
    (-0.5, -0.3, 0.0, 0.1, 0.2),
    (-0.2, -0.1, 0.1, 0.3, 0.4),
    (0.0, 0.2, 0.4, 0.6, 0.7),
    (0.1, 0.4, 0.7, 0.9, 1.1),
    (0.3, 0.6, 0.9, 1.2, 1.5),
)


def _listed(values):
    return field(default_factory=lambda: list(values))


def _all_cols(self) -> List[str]:
    """Raw columns needed from the CSV ('tenure' is derived from fiscalyear - ceo_year)."""
    parts = (self.CEO_CAT_COLS, self.FIRM_NUMERIC_COLS, self.FIRM_CAT_COLS, self.CEO_PROB_COLS, self.FIRM_PROB_COLS,
             ["Age", "fiscalyear", "ceo_year"])
    return list({c for part in parts for c in part})


StructuralConfig = make_dataclass(
    "StructuralConfig",
    [("DEVICE", torch.device, field(default_factory=pick_device))]
    + [(name, type(value), value) for name, value in schema.STRUCTURAL_DEFAULTS.items()]
    + [("BLM_INTERACTION_MATRIX", List[List[float]], field(default_factory=lambda: [list(r) for r in _BLM_A])),
       ("CEO_PROB_COLS", List[str], _listed(schema.posterior_columns("ceo"))),
       ("FIRM_PROB_COLS", List[str], _listed(schema.posterior_columns("firm"))),
       ("CEO_NUMERIC_COLS", List[str], _listed(("Age", "tenure"))),
       ("CEO_CAT_COLS", List[str], _listed(schema.CEO_CATEGORICAL)),
       ("FIRM_NUMERIC_COLS", List[str], _listed(schema.FIRM_NUMERIC)),
       ("FIRM_CAT_COLS", List[str], _listed(schema.FIRM_CATEGORICAL))],
    namespace={"all_cols": property(_all_cols), "__module__": __name__},
)
StructuralConfig.__doc__ = "Hyper-parameters and column names of the structural distillation run (see module docstring)."
