"""Training loop of the two-tower model — drop-in for the reference's ``training.py``.

``train_model(train_loader, val_loader, metadata, config)`` keeps the reference signature, prints and
returned object (``ceo_firm_matching/training.py:15-64``).  Underneath, each step is the fused CUDA path:
towers forward -> cosine head + weighted MSE -> head backward -> towers backward -> deterministic
embedding-gradient segment reduce, followed by the fused Adam step (``optim.FusedAdam``: torch.optim.Adam's state and
arithmetic in one launch).
Steps with a fixed batch shape are captured once into a CUDA graph and replayed (``GraphedTwoTowerStep``),
which removes the per-step Python / launch overhead that dominates at the reference's batch sizes; the
per-step ``loss.item()`` synchronisation of the reference is replaced by a device-side accumulator read once
per epoch (the printed values are the same).
"""
from typing import Dict, Optional, Sequence

import torch
import torch.optim as optim
from torch.utils.data import DataLoader

from . import ops
from .batching import device_batches
from .config import Config
from .model import CEOFirmMatcher

BATCH_KEYS = ("firm_numeric", "firm_cat", "ceo_numeric", "ceo_cat", "target", "weights")


class GraphedTwoTowerStep:
    """One training step (zero-grad, fwd, weighted-MSE loss, bwd[, optimiser]) for a fixed batch size,
    captured into a CUDA graph.  ``step(batch)`` copies the six batch tensors into static device buffers
    (host tensors are accepted: the copy is then the H2D transfer) and replays the graph; the returned loss
    is a device tensor that stays valid until the next call.  Dropout masks differ per replay (device-side
    Philox offset counter advanced by a graph node)."""

    def __init__(self, model: CEOFirmMatcher, example: Sequence[torch.Tensor],
                 optimizer: Optional[torch.optim.Optimizer] = None, warmup: int = 3,
                 stream: Optional[torch.cuda.Stream] = None, loss_scale: float = 1.0,
                 after_backward=None, before_forward=None):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTwoTowerStep needs the model on a CUDA device (no CPU fallback)")
        self.model, self.optimizer, self.device = model, optimizer, dev
        self.loss_scale = loss_scale             # data parallelism: 1/world so gradients are those of the global mean
        # multi-GPU: the gradient synchronisation (e.g. TableShardedTwoTower.sync_gradients) runs right after the
        # backward INSIDE the captured step, NCCL collectives included, so a step stays one graph launch per rank
        self.after_backward = after_backward
        self.before_forward = before_forward     # e.g. TableShardedTwoTower.begin_step(f_cat, c_cat)
        self.static = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in example]
        for s, t in zip(self.static, example):
            s.copy_(t)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        # seed of the backward pass (dL/dloss = loss_scale): passing it saves autograd's ones_like and the scaling multiply
        self._grad_seed = torch.full((), float(loss_scale), device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.loss = None
        # Autograd pins each parameter's gradient-accumulation node to the stream of its first use; capture fails
        # if that is the legacy default stream.  Warm-up, capture and replay therefore share one side stream (the
        # caller's own non-default stream when it already trains on one).
        self.stream = stream if stream is not None else torch.cuda.Stream(dev)
        if warmup == 0 and optimizer is not None and any(
                p.requires_grad and "exp_avg" not in optimizer.state.get(p, {}) for g in optimizer.param_groups for p in g["params"]):
            # the optimiser would create (zero-fill) its moments INSIDE the capture and every replay would reset them
            raise RuntimeError("GraphedTwoTowerStep(warmup=0) needs the optimiser state to exist: run one eager step of this "
                               "batch size first (train_model does), or pass warmup >= 1")
        ops.set_graph_rng_counter(self.counter)
        try:
            self.stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.stream):
                for _ in range(warmup):          # eager warm-up: allocations, func attributes, sort scratch
                    self._body()
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.loss = self._body()
        finally:
            ops.set_graph_rng_counter(None)

    def _body(self) -> torch.Tensor:
        ops.advance_graph_rng_counter()
        self.model.zero_grad_fast()
        if self.before_forward is not None:
            self.before_forward(self.static[1], self.static[3])
        loss, _ = self.model.forward_loss(*self.static)
        loss.backward(gradient=self._grad_seed)
        if self.after_backward is not None:
            self.after_backward()
        if self.optimizer is not None:
            self.optimizer.step()
            self.model.rezero_table_grads()      # tables are clean again before any other graph runs
        return loss.detach()

    def load(self, batch: Sequence[torch.Tensor]) -> None:
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)

    def step(self, batch: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        if batch is not None:
            self.load(batch)
        self.graph.replay()                      # replays on the caller's current stream
        return self.loss


def eager_step(model: CEOFirmMatcher, optimizer: Optional[torch.optim.Optimizer],
               batch: Sequence[torch.Tensor]) -> torch.Tensor:
    """The same step without graph capture (training.py:44-55)."""
    model.zero_grad_fast()
    loss, _ = model.forward_loss(*batch)
    loss.backward()
    if optimizer is not None:
        optimizer.step()
        model.rezero_table_grads()
    return loss.detach()


def _make_optimizer(model: CEOFirmMatcher, lr: float) -> optim.Optimizer:
    # optim.Adam(lr) of training.py:32 as ONE fused launch over all parameter tensors: bit-equal to
    # torch.optim.Adam(capturable=True), step counter on the device so the update lives inside the CUDA graph
    from .optim import FusedAdam
    return FusedAdam(model.parameters(), lr=lr)


def train_model(train_loader: DataLoader, val_loader: DataLoader, metadata: Dict[str, int],
                config: Config) -> Optional[CEOFirmMatcher]:
    """Train ``CEOFirmMatcher`` with Adam on the weighted MSE of ``match_means`` (training.py:15-64).

    ``val_loader`` is accepted and unused, exactly like the reference.
    """
    device = torch.device(config.DEVICE)
    if device.type != "cuda":
        raise RuntimeError("this build trains on CUDA only (config.DEVICE resolved to %s)" % device)
    model = CEOFirmMatcher(metadata, config).to(device)
    optimizer = _make_optimizer(model, config.LEARNING_RATE)
    model.use_persistent_table_grads(True)

    print(f"Starting training on {config.DEVICE} for {config.EPOCHS} epochs...")
    stream = torch.cuda.Stream(device)           # the whole loop runs on one side stream (see GraphedTwoTowerStep)
    stream.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(stream):
        _train_epochs(model, optimizer, train_loader, config, device, stream)
    torch.cuda.current_stream(device).wait_stream(stream)
    model.use_persistent_table_grads(False)
    return model


def _train_epochs(model, optimizer, train_loader, config, device, stream) -> None:
    # One captured step per distinct batch size.  The first step of a given size runs eagerly (it is a real
    # training step and leaves every buffer allocated); the second occurrence captures the graph.
    graphed: Dict[int, GraphedTwoTowerStep] = {}
    seen: Dict[int, int] = {}
    for epoch in range(config.EPOCHS):
        model.train()
        total_loss = torch.zeros((), device=device)
        n_batches = 0
        for tensors in device_batches(train_loader, BATCH_KEYS, device):
            B = tensors[0].shape[0]
            if B == 1:
                raise ValueError("Expected more than 1 value per channel when training")   # nn.BatchNorm1d
            seen[B] = seen.get(B, 0) + 1
            if seen[B] == 1:
                loss = eager_step(model, optimizer, tensors)
            else:
                if B not in graphed:
                    graphed[B] = GraphedTwoTowerStep(model, tensors, optimizer, warmup=0, stream=stream)
                loss = graphed[B].step(tensors)
            total_loss += loss
            n_batches += 1
        ops.raise_if_index_error(device)
        avg_loss = float(total_loss) / max(n_batches, 1)
        if epoch % 5 == 0:
            print(f"Epoch {epoch}: Avg Train Loss = {avg_loss:.4f}")
