"""Single-node multi-GPU partitioning of the hot path (one process per GPU, ``torch.distributed`` over NCCL /
NVLink; the reference has no distributed code at all — SURVEY.md 8e).

* ``DataParallelTwoTower``  – batch rows sharded over ranks, replicas kept bit-identical: one flat-buffer all-reduce
  for the dense tower parameters, and the embedding-table gradients rebuilt on every rank from the all-gathered
  (index, per-pair gradient row) pairs with the same deterministic sorted-segment reduce, in rank order.
* ``GlobalInfoNCE``         – in-batch InfoNCE with global negatives: projections are all-gathered (bf16), every
  rank runs the similarity-tile kernels for ITS rows against ALL columns (row sums in both directions, then both
  gradients), so no partial column sums or gradient reduce-scatter are needed; only two tiny all-gathers of the
  row sums and one scalar all-reduce.
* ``score_topk_sharded``    – CEO/firm row-sharded all-pairs scoring: column shards are all-gathered, every rank
  scores its rows against each shard (global column offsets) and the per-shard lists go through the k-way merge.

The communication protocol is independent of the kernels: ``backend`` objects provide the per-rank math (CUDA ops by
default); the CPU/gloo tests inject a torch reference backend to check the protocol end to end.
"""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------
# helpers (pure host logic, CPU-testable with gloo)
# ---------------------------------------------------------------------------------------------
def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced shard [lo, hi) of ``n`` rows for ``rank`` (first ``n % world`` ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_flat_(tensors: Sequence[torch.Tensor], group=None, scale: float = 1.0) -> None:
    """Sum ``tensors`` across ranks in ONE collective (they are a few tens of kB: latency-bound, so bucketing into a
    single flat buffer is what matters), optionally scale, and write the result back in place."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if scale != 1.0:
        flat.mul_(scale)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather equally sized row blocks ``[n, ...]`` into ``[world * n, ...]`` (rank order)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def gather_ragged_rows(x: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather row blocks of different lengths; returns the concatenation and the per-rank row counts."""
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    padded = torch.zeros((m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    padded[:x.shape[0]] = x
    allp = gather_rows(padded, group)
    parts = [allp[r * m:r * m + counts[r]] for r in range(world)]
    return torch.cat(parts), counts


# ---------------------------------------------------------------------------------------------
# data-parallel two-tower training
# ---------------------------------------------------------------------------------------------
class DataParallelTwoTower:
    """Keeps ``world`` replicas of a ``CEOFirmMatcher`` identical under data parallelism.

    Usage per step::

        model.zero_grad_fast()
        loss, _ = model.forward_loss(*local_batch)
        (loss * dp.loss_scale).backward()      # dp.loss_scale = 1 / world: mean over the GLOBAL batch
        dp.sync_gradients()                    # all ranks now hold the same gradients
        optimizer.step()

    BatchNorm uses the statistics of the LOCAL shard (the reference has no SyncBN), so the result
    equals the single-GPU run only per rank; replicas stay in sync because every gradient is synchronised.
    """

    def __init__(self, model, group=None):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.loss_scale = 1.0 / self.world
        self._table_ids = set()
        for h in model._handles:
            if h.table_grads is None:
                raise RuntimeError("call model.use_persistent_table_grads(True) before wrapping it for data parallelism")
            h.table_grads.defer = True             # backward stashes (indices, gradient rows) instead of reducing
            self._table_ids.update(id(e.weight) for e in h.embeddings)
        self.dense = [p for p in model.parameters() if id(p) not in self._table_ids]
        dist.barrier(group=group)
        for p in model.parameters():               # start from identical replicas
            dist.broadcast(p.data, src=0, group=group)
        for b in model.buffers():
            dist.broadcast(b.data, src=0, group=group)

    def sync_gradients(self) -> None:
        from . import ops
        allreduce_flat_([p.grad for p in self.dense], self.group)
        for h in self.model._handles:
            pg = h.table_grads
            if pg.pending is None:
                continue
            # not cleared: inside a replayed CUDA graph the same static buffers are refilled every step
            x_cat, dx_emb = pg.pending
            x_all = gather_rows(x_cat, self.group)                 # [world*B, K] int64, rank order
            dx_all = gather_rows(dx_emb, self.group)               # [world*B, K*E] f32
            ops.reduce_table_grads(h, x_all, dx_all)               # same inputs, same order -> bitwise equal replicas


# ---------------------------------------------------------------------------------------------
# InfoNCE with global negatives
# ---------------------------------------------------------------------------------------------
class CudaInfoNCEBackend:
    """Per-rank math of GlobalInfoNCE on the tcgen05 similarity-tile kernels."""

    def pack(self, x):
        from . import ops
        return ops.pack_bf16(x)

    def rowsum(self, xb, yb, temperature, diag_offset, want_diag):
        from . import ops
        return ops.infonce_rowsum(xb, yb, temperature, diag_offset, want_diag)

    def local_loss(self, rs_row, rs_col, diag, temperature, b_total):
        from . import _native as N
        loss = torch.empty((), device=rs_row.device)
        with torch.cuda.device(rs_row.device):
            N.check(N.lib().cfm_infonce_loss(N.ptr(rs_row), N.ptr(rs_col), N.ptr(diag), rs_row.shape[0], temperature,
                                             b_total, N.ptr(loss), N.stream_ptr()))
        return loss

    def grad(self, xb, yb, d, temperature, diag_offset, b_total, rs_x, rs_y, diag, g_loss):
        from . import ops
        return ops.infonce_grad(xb, yb, d, temperature, diag_offset, b_total, rs_x, rs_y, diag, g_loss)


class _GlobalInfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, firm_local, ceo_local, temperature, group, backend):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n, d = firm_local.shape
        b_total = n * world
        off = rank * n
        fb, cb = backend.pack(firm_local), backend.pack(ceo_local)
        fb_all, cb_all = gather_rows(fb, group), gather_rows(cb, group)
        rs_f, diag = backend.rowsum(fb, cb_all, temperature, off, True)        # row sums of S   for my firm rows
        rs_c, _ = backend.rowsum(cb, fb_all, temperature, off, False)          # row sums of S^T for my CEO rows
        loss = backend.local_loss(rs_f, rs_c, diag, temperature, b_total)
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)               # the global loss, equal on all ranks
        ctx.saved = (fb, cb, fb_all, cb_all, rs_f, rs_c, diag)
        ctx.meta = (temperature, d, off, b_total, group, backend, firm_local.dtype)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        fb, cb, fb_all, cb_all, rs_f, rs_c, diag = ctx.saved
        temperature, d, off, b_total, group, backend, dtype = ctx.meta
        rs_f_all, rs_c_all = gather_rows(rs_f, group), gather_rows(rs_c, group)
        g = g_loss.contiguous().float()
        d_firm = backend.grad(fb, cb_all, d, temperature, off, b_total, rs_f, rs_c_all, diag, g)
        d_ceo = backend.grad(cb, fb_all, d, temperature, off, b_total, rs_c, rs_f_all, diag, g)
        return d_firm.to(dtype), d_ceo.to(dtype), None, None, None


def info_nce_loss_global(firm_proj: torch.Tensor, ceo_proj: torch.Tensor, temperature: float = 0.07, group=None,
                         backend=None) -> torch.Tensor:
    """Symmetric InfoNCE (contrastive.py:102-138) over the GLOBAL batch: every rank passes its ``[B/world, D]`` row
    blocks (equal sizes), negatives are all rows of all ranks.  Returns the global loss (identical on every rank);
    its gradient w.r.t. the local rows is exact, so parameter gradients must be SUMMED across ranks."""
    backend = backend or CudaInfoNCEBackend()
    return _GlobalInfoNCEFn.apply(firm_proj, ceo_proj, float(temperature), group, backend)


# ---------------------------------------------------------------------------------------------
# row-sharded all-pairs scoring with cross-shard top-k merge
# ---------------------------------------------------------------------------------------------
class CudaScoringBackend:
    def topk(self, rows, cols, k, scale, col_offset):
        from .scoring import score_topk
        s, i, s64 = score_topk(rows, cols, k, scale, col_offset=col_offset, return_f64=True)
        return s64, i                                  # fp64 scores: the merge must order exactly

    def merge(self, part_scores, part_indices):
        from .scoring import merge_topk
        return merge_topk(part_scores, part_indices)


def score_topk_sharded(rows_local: torch.Tensor, cols_local: torch.Tensor, k: int, scale: float = 1.0, group=None,
                       backend=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-``k`` columns (global indices) for this rank's rows when BOTH sides are row-sharded across ranks:
    column shards are all-gathered (ragged shards allowed), each is scored separately with its global offset, and
    the ``world`` partial lists are k-way merged — the cross-GPU merge of BASELINE config 5."""
    backend = backend or CudaScoringBackend()
    cols_all, counts = gather_ragged_rows(cols_local.contiguous(), group)
    parts_s, parts_i = [], []
    off = 0
    for c in counts:
        if c > 0:
            s, i = backend.topk(rows_local, cols_all[off:off + c], k, scale, off)
            parts_s.append(s)
            parts_i.append(i)
        off += c
    if len(parts_s) == 1:
        return parts_s[0], parts_i[0]
    return backend.merge(torch.stack(parts_s), torch.stack(parts_i))
