"""ceo_firm_matching — B200-native build of the CEO-Recommender two-tower hot path.

Drop-in for the reference package's training/scoring surface: ``CEOFirmMatcher``,
``StructuralDistillationNet``, ``train_model``, ``train_structural_model``, ``contrastive``
(``ContrastiveCEOFirmMatcher``, ``info_nce_loss``, ``train_contrastive``, ``compute_retrieval_metrics``),
``scoring.score_topk``, ``encode.encode`` / ``encode.score_grid`` (the tower encoder for any module with the reference's
tower layout, batched heat-map scoring) and the ``cli`` / ``structural_cli`` ``--synthetic`` entry points.  The arithmetic
runs in hand-written sm_100a CUDA kernels (``libcfm_b200.so``) reached through a C ABI
(``include/cfm_b200.h``); there is no CPU fallback.  Plotting / explainability / WRDS modules of the
reference are out of scope.
"""
from .config import Config
from .structural_config import StructuralConfig
from .data import CEOFirmDataset, DataProcessor
from .model import CEOFirmMatcher
from .structural_data import DistillationDataset, StructuralDataProcessor
from .structural_model import StructuralDistillationNet
from .structural_training import train_structural_model
from .synthetic import generate_structural_synthetic_data, generate_synthetic_data
from .training import train_model

__version__ = "0.4.0+b200"

__all__ = [
    "Config", "DataProcessor", "CEOFirmDataset", "CEOFirmMatcher", "train_model",
    "generate_synthetic_data", "generate_structural_synthetic_data",
    "StructuralConfig", "StructuralDataProcessor", "DistillationDataset", "StructuralDistillationNet",
    "train_structural_model",
]
