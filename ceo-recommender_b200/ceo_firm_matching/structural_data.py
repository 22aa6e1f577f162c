"""Data preparation for the Structural Distillation Network — drop-in for the reference's
``structural_data.py`` (``ceo_firm_matching/structural_data.py:19-177``): loads the BLM-posterior CSV (or
falls back to the synthetic generator), derives tenure, renormalises the posteriors, splits 80/20
(``random_state=42``), fits encoders/scalers on the training part and returns tensor-dict datasets."""
import os
from typing import Any, Dict, Tuple

import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import train_test_split
from sklearn.preprocessing import LabelEncoder, StandardScaler
from torch.utils.data import Dataset

from .data import encode_known
from .structural_config import StructuralConfig

PROB_EPS = 1e-9


class DistillationDataset(Dataset):
    def __init__(self, data_dict: Dict[str, torch.Tensor]):
        self.data = data_dict
        self.length = len(data_dict["firm_num"])

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        return {k: v[idx] for k, v in self.data.items()}


class StructuralDataProcessor:
    def __init__(self, config: StructuralConfig):
        self.cfg = config
        self.encoders: Dict[str, LabelEncoder] = {}
        self.scalers: Dict[str, StandardScaler] = {"firm": StandardScaler(), "ceo": StandardScaler()}
        self.final_ceo_numeric = self.cfg.CEO_NUMERIC_COLS      # ['Age', 'tenure']
        self.final_firm_numeric = self.cfg.FIRM_NUMERIC_COLS

    @staticmethod
    def _add_tenure(df: pd.DataFrame) -> pd.DataFrame:
        if "tenure" not in df.columns and {"fiscalyear", "ceo_year"} <= set(df.columns):
            df["tenure"] = (df["fiscalyear"] - df["ceo_year"]).clip(lower=0)
        return df

    def _generate_synthetic(self, n: int = 2000) -> pd.DataFrame:
        from .synthetic import generate_structural_synthetic_data
        return generate_structural_synthetic_data(n_samples=n)

    def load_and_prep(self) -> Tuple[Dataset, Dataset, pd.DataFrame]:
        """Load -> engineer -> clean -> renormalise posteriors -> split -> fit -> tensorise."""
        if os.path.exists(self.cfg.DATA_PATH):
            df = pd.read_csv(self.cfg.DATA_PATH)
        else:
            print(f"Warning: Data not found at {self.cfg.DATA_PATH}. Generating SYNTHETIC data.")
            df = self._generate_synthetic()
        df = self._add_tenure(df)

        required = (list(self.cfg.CEO_CAT_COLS) + list(self.cfg.FIRM_NUMERIC_COLS) + list(self.cfg.FIRM_CAT_COLS)
                    + list(self.cfg.CEO_PROB_COLS) + list(self.cfg.FIRM_PROB_COLS))
        if "Age" in df.columns:
            required.append("Age")
        missing = [c for c in required if c not in df.columns]
        if missing:
            print(f"Warning: Missing columns: {missing}. Generating synthetic data instead.")
            df = self._add_tenure(self._generate_synthetic())
        df = df.dropna(subset=[c for c in required if c in df.columns]).reset_index(drop=True)

        # posteriors must be proper distributions for the KL loss (structural_data.py:74-78)
        for cols in (self.cfg.CEO_PROB_COLS, self.cfg.FIRM_PROB_COLS):
            p = df[cols].values
            df[cols] = p / (p.sum(axis=1, keepdims=True) + PROB_EPS)

        train_df, val_df = train_test_split(df, test_size=0.2, random_state=42)
        self._fit_transformers(train_df)
        return DistillationDataset(self._transform(train_df)), DistillationDataset(self._transform(val_df)), val_df

    def _fit_transformers(self, df: pd.DataFrame):
        for col in list(self.cfg.FIRM_CAT_COLS) + list(self.cfg.CEO_CAT_COLS):
            self.encoders[col] = LabelEncoder().fit(df[col].astype(str))
        self.scalers["firm"].fit(df[self.final_firm_numeric])
        self.scalers["ceo"].fit(df[self.final_ceo_numeric])

    def _transform(self, df: pd.DataFrame) -> Dict[str, torch.Tensor]:
        def cats(cols) -> torch.Tensor:
            codes = [encode_known(self.encoders[c], df[c].astype(str)) for c in cols]
            return torch.tensor(np.stack(codes, axis=1), dtype=torch.long)

        def f32(values) -> torch.Tensor:
            return torch.tensor(np.asarray(values), dtype=torch.float32)

        return {
            "firm_num": f32(self.scalers["firm"].transform(df[self.final_firm_numeric])),
            "firm_cat": cats(self.cfg.FIRM_CAT_COLS),
            "ceo_num": f32(self.scalers["ceo"].transform(df[self.final_ceo_numeric])),
            "ceo_cat": cats(self.cfg.CEO_CAT_COLS),
            "target_ceo": f32(df[self.cfg.CEO_PROB_COLS].values),
            "target_firm": f32(df[self.cfg.FIRM_PROB_COLS].values),
        }

    def get_metadata(self) -> Dict[str, Any]:
        return {
            "n_firm_num": len(self.final_firm_numeric),
            "n_ceo_num": len(self.final_ceo_numeric),
            "firm_cat_cards": [len(self.encoders[c].classes_) for c in self.cfg.FIRM_CAT_COLS],
            "ceo_cat_cards": [len(self.encoders[c].classes_) for c in self.cfg.CEO_CAT_COLS],
        }

    def get_feature_names(self) -> list:
        return (list(self.final_firm_numeric) + list(self.cfg.FIRM_CAT_COLS)
                + list(self.final_ceo_numeric) + list(self.cfg.CEO_CAT_COLS))
