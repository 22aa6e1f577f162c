"""Data preparation for the two-tower model — drop-in for the reference's ``data.py``.

CPU-side pandas / scikit-learn preprocessing that runs once per job (out of scope for kernels, SURVEY.md
section 2 row 8); it defines the input contract of the fused ops: ``firm_numeric [N,12] f32``,
``firm_cat [N,4] i64``, ``ceo_numeric [N,2] f32``, ``ceo_cat [N,7] i64``, ``target``/``weights [N,1] f32``
plus the metadata keys ``CEOFirmMatcher`` reads (reference ``ceo_firm_matching/data.py:16-198``).
"""
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd
import torch
from sklearn.preprocessing import LabelEncoder, StandardScaler
from torch.utils.data import Dataset

from .config import Config

WEIGHT_EPS = 1e-6


def encode_known(encoder: LabelEncoder, values: pd.Series) -> np.ndarray:
    """Label-encode ``values``; labels unseen at fit time map to code 0 (the reference's fallback, data.py:97-100).
    Vectorised: one dictionary lookup per row instead of one ``encoder.transform`` call per row."""
    lookup = {label: code for code, label in enumerate(encoder.classes_)}
    return values.map(lambda v: lookup.get(v, 0)).to_numpy(dtype=np.int64)


class DataProcessor:
    """Loads, cleans, engineers, encodes and scales; owns the fitted encoders / scalers."""

    def __init__(self, config: Config):
        self.cfg = config
        self.encoders: Dict[str, LabelEncoder] = {}
        self.scalers: Dict[str, StandardScaler] = {"firm": StandardScaler(), "ceo": StandardScaler()}
        self.final_firm_numeric = self.cfg.FIRM_NUMERIC_COLS
        self.final_ceo_numeric = self.cfg.CEO_NUMERIC_COLS + ["tenure"]
        self.processed_df: Optional[pd.DataFrame] = None

    # ---- I/O -------------------------------------------------------------------------------
    def load_data(self) -> pd.DataFrame:
        print(f"Loading data from {self.cfg.DATA_PATH}...")
        try:
            df = pd.read_csv(self.cfg.DATA_PATH, on_bad_lines="skip")
        except FileNotFoundError:
            print(f"Error: File not found at {self.cfg.DATA_PATH}")
            return pd.DataFrame()
        missing = [c for c in self.cfg.all_required_cols if c not in df.columns]
        if missing:
            print(f"Error: Missing essential columns: {missing}")
            return pd.DataFrame()
        return df[self.cfg.all_required_cols].copy()

    # ---- stateless feature engineering -------------------------------------------------------
    def prepare_features(self, df: pd.DataFrame) -> pd.DataFrame:
        if df.empty:
            return df
        n_before = len(df)
        df = df.dropna().copy()
        print(f"Dropped {n_before - len(df)} rows with NaNs. Final count: {len(df)}")
        df["tenure"] = (df["fiscalyear"] - df["ceo_year"]).clip(lower=0)
        df["weights"] = 1 / (df[self.cfg.WEIGHT_COL] ** 2 + WEIGHT_EPS)     # inverse-variance weights
        return df

    # ---- fit / transform ---------------------------------------------------------------------
    def fit(self, df: pd.DataFrame):
        print("Fitting scalers and encoders on training data...")
        for col in list(self.cfg.FIRM_CAT_COLS) + list(self.cfg.CEO_CAT_COLS):
            self.encoders[col] = LabelEncoder().fit(df[col].astype(str))
        self.scalers["firm"].fit(df[self.final_firm_numeric])
        self.scalers["ceo"].fit(df[self.final_ceo_numeric])

    def transform(self, df: pd.DataFrame) -> Dict[str, Any]:
        if df.empty:
            return {}
        df = df.copy()
        for col in list(self.cfg.FIRM_CAT_COLS) + list(self.cfg.CEO_CAT_COLS):
            df[f"{col}_code"] = encode_known(self.encoders[col], df[col].astype(str))
        df[self.final_firm_numeric] = self.scalers["firm"].transform(df[self.final_firm_numeric])
        df[self.final_ceo_numeric] = self.scalers["ceo"].transform(df[self.final_ceo_numeric])
        self.processed_df = df
        return self._to_tensors(df)

    def _to_tensors(self, df: pd.DataFrame) -> Dict[str, Any]:
        def cats(cols: List[str]) -> torch.Tensor:
            return torch.tensor(np.stack([df[f"{c}_code"].values for c in cols], axis=1), dtype=torch.long)

        def f32(values) -> torch.Tensor:
            return torch.tensor(np.asarray(values), dtype=torch.float32)

        out: Dict[str, Any] = {
            "firm_numeric": f32(df[self.final_firm_numeric].values),
            "firm_cat": cats(self.cfg.FIRM_CAT_COLS),
            "ceo_numeric": f32(df[self.final_ceo_numeric].values),
            "ceo_cat": cats(self.cfg.CEO_CAT_COLS),
            "target": f32(df[self.cfg.TARGET_COL].values).view(-1, 1),
            "weights": f32(df["weights"].values).view(-1, 1),
        }
        # the same dict doubles as the model metadata (cli.py:59 passes it to train_model)
        out.update({
            "n_firm_numeric": len(self.final_firm_numeric),
            "firm_cat_counts": [len(self.encoders[c].classes_) for c in self.cfg.FIRM_CAT_COLS],
            "n_ceo_numeric": len(self.final_ceo_numeric),
            "ceo_cat_counts": [len(self.encoders[c].classes_) for c in self.cfg.CEO_CAT_COLS],
        })
        return out

    def get_feature_names(self) -> List[str]:
        """Flat feature order used by wrappers: firm numeric, firm cat, CEO numeric, CEO cat."""
        return (list(self.final_firm_numeric) + list(self.cfg.FIRM_CAT_COLS)
                + list(self.final_ceo_numeric) + list(self.cfg.CEO_CAT_COLS))

    def get_flat_features(self, df: pd.DataFrame) -> np.ndarray:
        """``[firm numeric | firm cat | CEO numeric | CEO cat]`` as one 2-D array (for PDP/SHAP-style tools)."""
        d = self.transform(df)
        return np.hstack([d[k].numpy() for k in ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat")])


class CEOFirmDataset(Dataset):
    """Map-style dataset over the tensor dict (kept for API compatibility with the reference loaders)."""
    KEYS = ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat", "target", "weights")

    def __init__(self, data_dict: Dict[str, Any]):
        self.data = data_dict
        self.length = len(data_dict["target"])

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        return {k: self.data[k][idx] for k in self.KEYS}
