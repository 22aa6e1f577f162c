"""Training loop of the Structural Distillation Network — drop-in for the reference's
``structural_training.py`` (``ceo_firm_matching/structural_training.py:17-117``): Adam on the sum of the two
KL(batchmean) distillation losses, a validation pass per epoch, best-validation tracking and the same prints.
Each step runs the fused towers and the fused softmax.A.softmax + KL head (loss and d/dlogits in one launch).
"""
from typing import Optional

import torch
import torch.optim as optim
from torch.utils.data import DataLoader

from . import ops
from .batching import device_batches
from .structural_config import StructuralConfig
from .structural_model import StructuralDistillationNet

BATCH_KEYS = ("firm_num", "firm_cat", "ceo_num", "ceo_cat", "target_ceo", "target_firm")


def _loss(model: StructuralDistillationNet, batch) -> torch.Tensor:
    f_num, f_cat, c_num, c_cat, t_ceo, t_firm = batch
    c_logits, f_logits = model.logits(f_num, f_cat, c_num, c_cat)
    return model.distillation_loss(c_logits, f_logits, t_ceo, t_firm)


class GraphedStructuralStep:
    """One distillation step (zero-grad, towers, KL head, backward, Adam) for a fixed batch size, captured into a CUDA
    graph: at the shipped batch size (128, ``structural_cli.py:87-97``) the ~40 launches of a step are pure launch
    latency, so a replay is several times faster than re-issuing them.  Dropout masks differ per replay (device-side
    Philox offset counter advanced by a graph node), exactly as in ``training.GraphedTwoTowerStep``."""

    def __init__(self, model: StructuralDistillationNet, optimizer, example, stream: torch.cuda.Stream):
        dev = example[0].device
        self.model, self.optimizer, self.stream = model, optimizer, stream
        self.static = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in example]
        for s, t in zip(self.static, example):
            s.copy_(t)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        ops.set_graph_rng_counter(self.counter)
        # No eager warm-up here: it would be an extra optimiser step.  The caller has already run one eager step of
        # this batch size (optimiser state, scratch buffers and kernel attributes exist), as train_model does.
        try:
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(self.graph, stream=stream):
                self.loss = self._body()
        finally:
            ops.set_graph_rng_counter(None)

    def _body(self) -> torch.Tensor:
        ops.advance_graph_rng_counter()
        self.optimizer.zero_grad(set_to_none=True)
        loss = _loss(self.model, self.static)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def step(self, batch) -> torch.Tensor:
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss


def train_structural_model(train_loader: DataLoader, val_loader: DataLoader, metadata: dict,
                           config: StructuralConfig) -> Optional[StructuralDistillationNet]:
    device = torch.device(config.DEVICE)
    if device.type != "cuda":
        raise RuntimeError("this build trains on CUDA only (config.DEVICE resolved to %s)" % device)
    model = StructuralDistillationNet(metadata, config).to(device)

    print("Initialized Structural Distillation Network.")
    print(f"  Device: {config.DEVICE}")
    print(f"  Interaction Matrix Frozen: {not model.A.requires_grad}")
    print(f"  CEO Tower Input: {metadata['n_ceo_num']} numeric + {len(metadata['ceo_cat_cards'])} categorical")
    print(f"  Firm Tower Input: {metadata['n_firm_num']} numeric + {len(metadata['firm_cat_cards'])} categorical")

    from .optim import FusedAdam
    optimizer = FusedAdam(model.parameters(), lr=config.LEARNING_RATE)   # optim.Adam's state/arithmetic, one launch
    print(f"\nStarting Distillation Training for {config.EPOCHS} epochs...")

    best_val_loss = float("inf")
    # One captured step per distinct train batch size (the first step of a size runs eagerly and leaves the optimiser
    # state allocated; the loop runs on one side stream, see training.GraphedTwoTowerStep).
    stream = torch.cuda.Stream(device)
    stream.wait_stream(torch.cuda.current_stream(device))
    graphed, seen = {}, {}
    for epoch in range(config.EPOCHS):
        model.train()
        train_loss = torch.zeros((), device=device)       # accumulated on the device: one sync per epoch
        n_train = 0
        with torch.cuda.stream(stream):
            for batch in device_batches(train_loader, BATCH_KEYS, device):
                B = batch[0].shape[0]
                seen[B] = seen.get(B, 0) + 1
                if seen[B] == 1:
                    optimizer.zero_grad(set_to_none=True)
                    loss = _loss(model, batch)
                    loss.backward()
                    optimizer.step()
                    loss = loss.detach()
                else:
                    if B not in graphed:
                        graphed[B] = GraphedStructuralStep(model, optimizer, batch, stream)
                    loss = graphed[B].step(batch)
                train_loss += loss
                n_train += 1
        torch.cuda.current_stream(device).wait_stream(stream)

        model.eval()
        val_loss = torch.zeros((), device=device)
        n_val = 0
        with torch.no_grad():
            for batch in device_batches(val_loader, BATCH_KEYS, device):
                val_loss += _loss(model, batch)
                n_val += 1
        ops.raise_if_index_error(device)
        avg_train_loss = float(train_loss) / max(n_train, 1)
        avg_val_loss = float(val_loss) / max(n_val, 1)
        best_val_loss = min(best_val_loss, avg_val_loss)
        if epoch % 10 == 0 or epoch == config.EPOCHS - 1:
            print(f"  Epoch {epoch:3d}: Train Loss = {avg_train_loss:.4f}, Val Loss = {avg_val_loss:.4f}")

    print(f"\nTraining Complete. Best Validation Loss: {best_val_loss:.4f}")
    return model
