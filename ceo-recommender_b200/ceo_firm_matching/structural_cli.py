"""``python -m ceo_firm_matching.structural_cli --synthetic --epochs E --batch-size B`` — drop-in for the
reference's structural CLI (``ceo_firm_matching/structural_cli.py:16-148``) up to the trained model; the
IlluminationEngine plotting tail needs seaborn/matplotlib (absent, out of scope) and is replaced by a
type-probability summary computed with the fused eval path."""
import argparse
import os

import torch
from torch.utils.data import DataLoader

from .structural_config import StructuralConfig
from .structural_data import StructuralDataProcessor
from .structural_training import train_structural_model


# (flag, StructuralConfig field it overrides, type, help) — the reference's command line, structural_cli.py:18-31
_OVERRIDES = (
    ("--epochs", "EPOCHS", int, "Number of training epochs (default: 50)"),
    ("--batch-size", "BATCH_SIZE", int, "Batch size for training (default: 256)"),
    ("--data-path", "DATA_PATH", str, "Path to BLM posteriors CSV file"),
    ("--output-path", "OUTPUT_PATH", str, "Output directory for results"),
)


def _config_from(argv) -> StructuralConfig:
    parser = argparse.ArgumentParser(description="Train Structural Distillation Network with BLM Priors")
    parser.add_argument("--synthetic", action="store_true", help="Use synthetic data for verification")
    for flag, _, kind, text in _OVERRIDES:
        parser.add_argument(flag, type=kind, default=None, help=text)
    args = parser.parse_args(argv)
    config = StructuralConfig()
    for flag, name, _, _ in _OVERRIDES:
        value = getattr(args, flag.lstrip("-").replace("-", "_"))
        if value:
            setattr(config, name, value)
    if args.synthetic:
        config.DATA_PATH = "SYNTHETIC_MODE"          # a path that does not exist forces the generator
    return config


def main(argv=None):
    config = _config_from(argv)
    banner = "=" * 60
    print(f"{banner}\nSTRUCTURAL DISTILLATION NETWORK\n{banner}")
    print(f"Device: {config.DEVICE}  Epochs: {config.EPOCHS}  Batch Size: {config.BATCH_SIZE}")

    processor = StructuralDataProcessor(config)
    train_ds, val_ds, val_df = processor.load_and_prep()
    print(f"Train size: {len(train_ds)}, Val size: {len(val_ds)}")
    train_loader = DataLoader(train_ds, batch_size=config.BATCH_SIZE, shuffle=True, drop_last=True)
    val_loader = DataLoader(val_ds, batch_size=config.BATCH_SIZE, shuffle=False)

    model = train_structural_model(train_loader, val_loader, processor.get_metadata(), config)
    if model is None:
        print("Error: Training failed!")
        return 1

    model.eval()
    with torch.no_grad():
        d = val_ds.data
        dev = config.DEVICE
        c_logits, f_logits, match = model(d["firm_num"].to(dev), d["firm_cat"].to(dev), d["ceo_num"].to(dev),
                                          d["ceo_cat"].to(dev))
        print(f"Validation expected match: mean {float(match.mean()):.4f}, std {float(match.std()):.4f}")
        print("Mean CEO type probabilities:", [round(x, 3) for x in torch.softmax(c_logits, 1).mean(0).tolist()])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
