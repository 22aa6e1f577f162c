"""Synthetic CEO-firm panels for ``--synthetic`` runs — drop-in for the reference's ``synthetic.py``.

The reference draws every column from the global NumPy stream right after ``np.random.seed(42)``
(``ceo_firm_matching/synthetic.py:10-75``), so the *order* of the draws is part of the contract: the
column spec below lists them in that order and reproduces the reference frames value for value
(checked by tests/test_synthetic_data.py against the reference when it is present, and against a
committed checksum otherwise).  Scaled generators for the large benchmark shapes live in ``bench.py``.
"""
import numpy as np
import pandas as pd

# (column, draw kind, arguments) in the reference's draw order
_COLUMNS = [
    ("gvkey", "randint", (1000, 9999)),
    ("match_exec_id", "randint", (10000, 99999)),
    ("Age", "uniform", (30, 70)),
    ("Output", "randint", (0, 2)),
    ("Throghput", "randint", (0, 2)),          # the reference's spelling, part of the schema
    ("Peripheral", "randint", (0, 2)),
    ("Gender", "choice", (["M", "F"],)),
    ("maxedu", "randint", (1, 5)),
    ("ivy", "randint", (0, 2)),
    ("m", "randint", (0, 2)),
    ("ceo_year", "randint", (2000, 2023)),
    ("year_born", "randint", (1950, 1990)),
    ("dep_baby_ceo", "randint", (0, 2)),
    ("DOB", "dates", ("1950-01-01",)),
    ("ind_firms_60w", "normal", (0, 1)),
    ("non_competition_score", "uniform", (0, 1)),
    ("boardindpw", "uniform", (0, 1)),
    ("boardsizew", "randint", (5, 20)),
    ("busyw", "randint", (0, 5)),
    ("pct_blockw", "uniform", (0, 100)),
    ("logatw", "uniform", (5, 15)),
    ("exp_roa", "normal", (0.05, 0.02)),
    ("rdintw", "uniform", (0, 0.2)),
    ("capintw", "uniform", (0, 0.3)),
    ("leverage", "uniform", (0, 1)),
    ("divyieldw", "uniform", (0, 0.05)),
    ("compindustry", "choice", (["Tech", "Finance", "Health", "Energy"],)),
    ("ba_state", "choice", (["CA", "NY", "TX", "MA"],)),
    ("rd_control", "randint", (0, 2)),
    ("dpayer", "randint", (0, 2)),
    ("fiscalyear", "randint", (2000, 2023)),
    ("match_means", "normal", (0, 1)),
    ("sd_match_means", "uniform", (0.1, 1.0)),
    ("mover", "randint", (0, 2)),               # legacy columns kept for schema compatibility
    ("output_exp_dummy", "randint", (0, 2)),
]


def _draw(kind: str, args: tuple, n: int):
    if kind == "randint":
        return np.random.randint(args[0], args[1], n)
    if kind == "uniform":
        return np.random.uniform(args[0], args[1], n)
    if kind == "normal":
        return np.random.normal(args[0], args[1], n)
    if kind == "choice":
        return np.random.choice(args[0], n)
    if kind == "dates":
        return pd.date_range(start=args[0], periods=n).strftime("%Y-%m-%d")
    raise ValueError(kind)


def generate_synthetic_data(n_samples: int = 1000) -> pd.DataFrame:
    """Synthetic frame with the two-tower schema (seed 42, like the reference)."""
    np.random.seed(42)
    return pd.DataFrame({name: _draw(kind, args, n_samples) for name, kind, args in _COLUMNS})


def generate_structural_synthetic_data(n_samples: int = 2000, seed: int = 42) -> pd.DataFrame:
    """Base synthetic frame plus Dirichlet(1,..,1) BLM posteriors ``prob_ceo_1..5`` / ``prob_firm_1..5``
    and the derived ``tenure`` column (reference synthetic.py:78-111)."""
    np.random.seed(seed)
    df = generate_synthetic_data(n_samples)      # re-seeds with 42 internally, like the reference
    for side in ("ceo", "firm"):
        probs = np.random.dirichlet(np.ones(5), n_samples)
        for k in range(5):
            df[f"prob_{side}_{k + 1}"] = probs[:, k]
    df["tenure"] = (df["fiscalyear"] - df["ceo_year"]).clip(lower=0)
    return df
