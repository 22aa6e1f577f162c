"""Configuration of the two-tower recommender — drop-in for the reference's ``config.py``.

Same attribute names and values as ``ceo_firm_matching/config.py:10-55`` of the reference; the only
behavioural difference is the device pick: this build runs on CUDA only (B200), so ``DEVICE`` is
``cuda`` whenever a GPU is visible.  Without a GPU it reports ``cpu`` so that data preparation and
module construction still work, but every model forward raises (there is no CPU fallback).
"""
from typing import List

import torch


def pick_device() -> torch.device:
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


class Config:
    # --- system ---
    DEVICE = pick_device()
    DATA_PATH = "Data/ceo_types_v0.2.csv"
    OUTPUT_PATH = "./Output"

    # --- optimisation ---
    EPOCHS = 40
    LEARNING_RATE = 0.0004
    LATENT_DIM = 60
    BATCH_SIZE = 128

    # --- embedding widths ---
    EMBEDDING_DIM_SMALL = 2
    EMBEDDING_DIM_MEDIUM = 8
    EMBEDDING_DIM_LARGE = 48

    # --- columns ---
    ID_COLS = ["gvkey", "match_exec_id"]

    CEO_NUMERIC_COLS = ["Age"]                      # 'tenure' is derived in the data processor
    CEO_CAT_COLS = ["Gender", "maxedu", "ivy", "m", "Output", "Throghput", "Peripheral"]
    CEO_RAW_COLS = CEO_NUMERIC_COLS + CEO_CAT_COLS + ["ceo_year", "dep_baby_ceo"]

    FIRM_NUMERIC_COLS = [
        "ind_firms_60w", "non_competition_score", "boardindpw", "boardsizew", "busyw", "pct_blockw",
        "logatw", "exp_roa", "rdintw", "capintw", "leverage", "divyieldw",
    ]
    FIRM_CAT_COLS = ["compindustry", "ba_state", "rd_control", "dpayer"]
    FIRM_RAW_COLS = FIRM_NUMERIC_COLS + FIRM_CAT_COLS + ["fiscalyear"]

    TARGET_COL = "match_means"
    WEIGHT_COL = "sd_match_means"

    @property
    def all_required_cols(self) -> List[str]:
        """Every distinct raw column the CSV must provide."""
        needed = self.ID_COLS + self.CEO_RAW_COLS + self.FIRM_RAW_COLS + [self.TARGET_COL, self.WEIGHT_COL]
        return list(set(needed))
