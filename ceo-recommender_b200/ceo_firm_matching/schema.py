"""Column vocabulary of the two CSV inputs (match panel, BLM posteriors), shared by both configurations.

The names are data facts of the reference's input files (config.py:33-47, structural_config.py:50-66); keeping them in
one table makes the two config classes views of the same schema."""
from typing import Dict, Tuple

FIRM_NUMERIC: Tuple[str, ...] = (
    "ind_firms_60w", "non_competition_score", "boardindpw", "boardsizew", "busyw", "pct_blockw",
    "logatw", "exp_roa", "rdintw", "capintw", "leverage", "divyieldw",
)
FIRM_CATEGORICAL: Tuple[str, ...] = ("compindustry", "ba_state", "rd_control", "dpayer")
CEO_CATEGORICAL: Tuple[str, ...] = ("Gender", "maxedu", "ivy", "m", "Output", "Throghput", "Peripheral")
KEYS: Tuple[str, ...] = ("gvkey", "match_exec_id")

# scalar defaults of the two-tower run (config.py:15-27) and of the structural run (structural_config.py:24-36)
TWO_TOWER_DEFAULTS: Dict[str, object] = dict(
    DATA_PATH="Data/ceo_types_v0.2.csv", OUTPUT_PATH="./Output",
    EPOCHS=40, LEARNING_RATE=0.0004, LATENT_DIM=60, BATCH_SIZE=128,
    EMBEDDING_DIM_SMALL=2, EMBEDDING_DIM_MEDIUM=8, EMBEDDING_DIM_LARGE=48,
    TARGET_COL="match_means", WEIGHT_COL="sd_match_means",
)
STRUCTURAL_DEFAULTS: Dict[str, object] = dict(
    DATA_PATH="Data/blm_posteriors.csv", OUTPUT_PATH="./Output/Structural_Distillation",
    EPOCHS=50, LEARNING_RATE=0.001, BATCH_SIZE=256, DROPOUT=0.2, LATENT_DIM=128, EMBEDDING_DIM=8,
)


def posterior_columns(side: str, n_types: int = 5) -> Tuple[str, ...]:
    return tuple(f"prob_{side}_{k}" for k in range(1, n_types + 1))
