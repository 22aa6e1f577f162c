"""Configuration of the two-tower recommender — drop-in for the reference's ``config.py``.

Same attribute names and values as ``ceo_firm_matching/config.py:10-55`` of the reference (they are assembled from
``schema.py``); the only behavioural difference is the device pick: this build runs on CUDA only (B200), so ``DEVICE``
is ``cuda`` whenever a GPU is visible.  Without a GPU it reports ``cpu`` so that data preparation and module
construction still work, but every model forward raises (there is no CPU fallback).
"""
from typing import List

import torch

from . import schema


def pick_device() -> torch.device:
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


class Config:
    DEVICE = pick_device()
    vars().update(schema.TWO_TOWER_DEFAULTS)          # EPOCHS, LEARNING_RATE, LATENT_DIM, EMBEDDING_DIM_*, ...

    ID_COLS = list(schema.KEYS)
    FIRM_NUMERIC_COLS = list(schema.FIRM_NUMERIC)
    FIRM_CAT_COLS = list(schema.FIRM_CATEGORICAL)
    CEO_NUMERIC_COLS = ["Age"]                        # 'tenure' is derived in the data processor
    CEO_CAT_COLS = list(schema.CEO_CATEGORICAL)
    FIRM_RAW_COLS = FIRM_NUMERIC_COLS + FIRM_CAT_COLS + ["fiscalyear"]
    CEO_RAW_COLS = CEO_NUMERIC_COLS + CEO_CAT_COLS + ["ceo_year", "dep_baby_ceo"]

    @property
    def all_required_cols(self) -> List[str]:
        """Every distinct raw column the CSV must provide."""
        wanted = [*self.ID_COLS, *self.CEO_RAW_COLS, *self.FIRM_RAW_COLS, self.TARGET_COL, self.WEIGHT_COL]
        return list(set(wanted))
