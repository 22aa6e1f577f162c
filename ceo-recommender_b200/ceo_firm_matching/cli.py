"""``python -m ceo_firm_matching.cli --synthetic`` — drop-in for the reference's two-tower CLI
(``ceo_firm_matching/cli.py:17-89``): data -> 80/20 split -> fit -> loaders (batch 256) -> ``train_model``.
The reference's PDP / heat-map tail needs matplotlib (absent in this image, plotting is out of scope); the CLI
instead reports the validation weighted-MSE of the trained model through the fused eval path."""
import argparse

import torch
from sklearn.model_selection import train_test_split
from torch.utils.data import DataLoader

from .config import Config
from .data import CEOFirmDataset, DataProcessor
from .training import train_model


def evaluate(model, data, device) -> float:
    """Weighted MSE (training.py:52) of ``model`` on a transformed data dict, eval mode."""
    model.eval()
    with torch.no_grad():
        ins = [data[k].to(device) for k in ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat")]
        preds = model(*ins)
        return float((data["weights"].to(device) * (preds - data["target"].to(device)) ** 2).mean())


def main(argv=None):
    parser = argparse.ArgumentParser(description="Train Two Towers Model")
    parser.add_argument("--synthetic", action="store_true", help="Use synthetic data for verification")
    args = parser.parse_args(argv)

    config = Config()
    print(f"Running Two Towers Model on {config.DEVICE}")
    processor = DataProcessor(config)
    if args.synthetic:
        print("Using SYNTHETIC data...")
        from .synthetic import generate_synthetic_data
        raw_df = generate_synthetic_data(1000)
    else:
        raw_df = processor.load_data()
    if raw_df.empty:
        return None

    df_clean = processor.prepare_features(raw_df)
    train_df, val_df = train_test_split(df_clean, test_size=0.2, random_state=42)
    print(f"Train size: {len(train_df)}, Val size: {len(val_df)}")
    processor.fit(train_df)
    train_data, val_data = processor.transform(train_df), processor.transform(val_df)
    train_loader = DataLoader(CEOFirmDataset(train_data), batch_size=256, shuffle=True)
    val_loader = DataLoader(CEOFirmDataset(val_data), batch_size=256, shuffle=False)

    model = train_model(train_loader, val_loader, train_data, config)
    if model is not None:
        print(f"Validation weighted MSE: {evaluate(model, val_data, config.DEVICE):.4f}")
    return model


if __name__ == "__main__":
    main()
