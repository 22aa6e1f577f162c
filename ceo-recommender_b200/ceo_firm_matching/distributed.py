"""Single-node multi-GPU partitioning of the hot path (one process per GPU, ``torch.distributed`` over NCCL /
NVLink; the reference has no distributed code at all — SURVEY.md 8e).

* ``DataParallelTwoTower``  – batch rows sharded over ranks, replicas kept bit-identical: one flat-buffer all-reduce
  for the dense tower parameters, and the embedding-table gradients rebuilt on every rank from the all-gathered
  (index, per-pair gradient row) pairs with the same deterministic sorted-segment reduce, in rank order.
* ``TableShardedTwoTower``  – the scalable form of the same step: every embedding table lives on ONE rank, the
  others map it over NVLink (CUDA IPC).  Rows are pulled peer-to-peer into a local stash, the towers stay data
  parallel, and each table's owner runs the sorted-segment reduce over every rank's gradient rows, read in place from
  the peers' buffers.  Per-rank exchange volume is independent of the world size (the replicated form above receives
  ``world`` times the gradient rows on every rank).
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
    views, off = [], 0
    for t in tensors:
        n = t.numel()
        views.append(flat[off:off + n].view_as(t))
        off += n
    torch._foreach_copy_(list(tensors), views)       # one multi-tensor kernel instead of one copy per parameter


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
# table-sharded two-tower training over NVLink peer memory
# ---------------------------------------------------------------------------------------------
def plan_table_slices(towers: Sequence[Tuple[int, int]], world: int, max_pieces: int = 4,
                      slack: float = 0.05) -> Tuple[List[int], List[Tuple[int, int, int]], List[int]]:
    """Decide which rank owns what.  ``towers`` = [(n_tables, emb_dim), ...].  Every table is cut into ``pieces``
    equal column slices (per tower; slice width a multiple of 4 floats so rows stay 16-byte vectors) and the
    slices are spread over the ranks longest-first onto the least loaded rank, load = slice width (bytes moved per
    pair).  The smallest ``pieces`` whose worst rank is within ``slack`` of the best achievable plan is kept.
    Returns (pieces per tower, slices [(tower, table, piece)], owner rank per slice) — identical on every rank."""
    def options(E):
        return [S for S in range(1, max_pieces + 1) if E % S == 0 and (E // S) % 4 == 0] or [1]

    def spread(pieces):
        slices = [(t, k, c) for t, (K, E) in enumerate(towers) for k in range(K) for c in range(pieces[t])]
        weight = [towers[t][1] // pieces[t] for t, _, _ in slices]
        load = [0] * world
        owner = [0] * len(slices)
        for s in sorted(range(len(slices)), key=lambda s: (-weight[s], s)):
            r = min(range(world), key=lambda r: (load[r], r))
            owner[s] = r
            load[r] += weight[s]
        return slices, owner, max(load)

    import itertools
    plans = []
    for pieces in itertools.product(*[options(E) for _, E in towers]):
        slices, owner, worst = spread(pieces)
        plans.append((worst, sum(pieces), list(pieces), slices, owner))
    best = min(p[0] for p in plans)
    ok = [p for p in plans if p[0] <= best * (1.0 + slack)]
    _, _, pieces, slices, owner = min(ok, key=lambda p: (p[1], p[0]))
    return pieces, slices, owner


class CudaIpcPeers:
    """Maps buffers of the other ranks of the node into this process: every rank exports the cudaMalloc allocation
    behind each published tensor (``cfm_ipc_export``), the 64-byte handles travel through the process group, and
    each accessor opens them with its own device current (``cfm_ipc_open``: NVLink peer access is enabled lazily)."""

    def __init__(self, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._opened = {}                          # (rank, handle) -> mapped base address (a handle opens once)

    def share(self, named: dict) -> dict:
        """name -> per-rank list (None where a rank did not publish that name); own entries are the local
        tensors, the others ``PeerView``s."""
        import ctypes as C
        from . import _native as N
        payload = {}
        for k, v in named.items():
            if not (v.is_cuda and v.is_contiguous()):
                raise ValueError(f"{k}: only contiguous CUDA tensors can be published")
            handle, offset = C.create_string_buffer(64), N.i64()
            N.check(N.lib().cfm_ipc_export(N.ptr(v), handle, C.byref(offset)))
            payload[k] = (handle.raw, offset.value, tuple(v.shape), v.dtype)
        gathered: List[Optional[dict]] = [None] * self.world
        dist.all_gather_object(gathered, payload, group=self.group)
        out = {}
        for r in range(self.world):
            for k, (handle, offset, shape, dtype) in gathered[r].items():
                if r == self.rank:
                    view = named[k]
                else:
                    if (r, handle) not in self._opened:
                        base = C.c_void_p()
                        N.check(N.lib().cfm_ipc_open(handle, C.byref(base)))
                        self._opened[(r, handle)] = base.value
                    view = N.PeerView(self._opened[(r, handle)] + offset, shape, dtype, r)
                out.setdefault(k, [None] * self.world)[r] = view
        return out

    def fence(self) -> None:
        """Called by every rank between the barrier collective and the owners' reads: peer memory is coherent at
        kernel boundaries, nothing to do (an emulation without shared memory refreshes its copies here)."""

    def close(self) -> None:
        from . import _native as N
        for base in self._opened.values():
            N.lib().cfm_ipc_close(base)
        self._opened.clear()


class CudaTableOps:
    """Per-rank kernels of the table-sharded step."""

    def make_row_source(self, handle, tables, pieces, dx_emb):
        from . import ops
        return ops.StashedRows(handle, tables, pieces, dx_emb)

    def make_scratch(self, n_owned, n_peers, B, device):
        from . import ops
        return ops._SortScratch(n_owned * n_peers * B, n_owned * n_peers, B, device)

    @staticmethod
    def _groups(groups, n_peers):
        from . import _native as N
        keep = []                                    # ctypes arrays referenced by pointer must outlive the call
        arr = (N.PeerGroup * len(groups))()
        for g, grp in enumerate(groups):
            owned = (N.PeerTable * len(grp["owned"]))()
            for j, o in enumerate(grp["owned"]):
                owned[j].n_cols, owned[j].col, owned[j].col0 = o["n_cols"], o["col"], o["col0"]
                owned[j].rows, owned[j].grad = o["grad"].shape[0], N.ptr(o["grad"])
                for r in range(n_peers):
                    owned[j].x_cat[r] = N.ptr(o["x_cat"][r])
                    owned[j].dx_emb[r] = N.ptr(o["dx_emb"][r])
            keep.append(owned)
            arr[g].owned, arr[g].n_owned = owned, len(grp["owned"])
            arr[g].emb_dim, arr[g].width = grp["emb_dim"], grp["width"]
        return arr, keep

    def peer_reduce(self, groups, n_peers, B, scratch, phase=0):
        """phase 0: keys + sort + reduce; 1: keys + sort (needs only the peers' indices); 2: reduce."""
        from . import _native as N
        arr, keep = self._groups(groups, n_peers)
        N.check(N.lib().cfm_emb_grad_peer_reduce(
            arr, len(groups), n_peers, B, phase, N.ptr(scratch.keys_tmp), N.ptr(scratch.vals_tmp),
            N.ptr(scratch.keys_sorted), N.ptr(scratch.vals_sorted), N.ptr(scratch.tmp), scratch.tmp_bytes,
            N.stream_ptr()))
        del keep

    def rezero(self, groups, n_peers, B, scratch):
        from . import _native as N
        arr, keep = self._groups(groups, 0)
        N.check(N.lib().cfm_emb_grad_peer_rezero(arr, len(groups), n_peers, B, N.ptr(scratch.keys_sorted),
                                                 N.stream_ptr()))
        del keep

    def side_stream(self, device):
        # high priority: the short sort kernels slot in between the CTAs of the long tower kernels
        return torch.cuda.Stream(device, priority=-1)


class TableShardedTwoTower:
    """Hybrid parallelism for ``CEOFirmMatcher`` on one NVLink node: towers data-parallel, tables sharded.

    * every embedding table (or column slice of a wide table, ``plan_table_slices``) is owned by one rank; only the
      owner's copy is read, receives gradients and should be stepped by the optimiser (``owned_parameters()``);
      ``consolidate()`` broadcasts the owners' slices so that ``state_dict()`` is complete on every rank again;
    * forward: the rows of the local batch are pulled from the owners over NVLink into a local stash
      (``cfm_emb_gather_rows``) that both the forward and the stage-1 backward read;
    * backward: every rank leaves its per-pair gradient rows in a fixed peer-visible buffer; after the dense
      all-reduce (which is also the barrier that makes those buffers complete) each owner runs the sorted-segment
      reduce over all ranks' rows, read in place through its peer mappings; a second tiny collective releases the
      buffers for the next step.

    Usage per step (``batch_rows`` rows per rank, fixed)::

        model.zero_grad_fast()
        loss, _ = model.forward_loss(*local_batch)
        (loss * ts.loss_scale).backward()
        ts.sync_gradients(release=False)
        optimizer.step()          # optimiser built over ts.owned_parameters()
        ts.release()              # peers may read the updated tables / overwrite their buffers from here on

    (without an optimiser ``ts.sync_gradients()`` alone does both).  The owner's dense table gradient is bitwise
    what ``DataParallelTwoTower`` produces on every rank: same items, same (rank, row) order inside each run."""

    def __init__(self, model, batch_rows: int, group=None, peers=None, table_ops=None):
        self.model, self.group, self.B = model, group, int(batch_rows)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.loss_scale = 1.0 / self.world
        self.kernels = table_ops or CudaTableOps()
        handles = self.handles = model._handles
        dev = next(model.parameters()).device
        self.pieces, self.slices, self.owner = plan_table_slices([(h.n_tables, h.emb_dim) for h in handles],
                                                                 self.world)
        table_ids = set()
        for h in handles:
            if h.table_grads is None:
                raise RuntimeError("call model.use_persistent_table_grads(True) before sharding the tables")
            h.table_grads.defer = True
            h.table_grads.rezero_hook = lambda: None      # the owner re-zeroes inside sync_gradients
            table_ids.update(id(e.weight) for e in h.embeddings)
        self.dense = [p for p in model.parameters() if id(p) not in table_ids]
        dist.barrier(group=group)
        for p in self.dense:
            dist.broadcast(p.data, src=0, group=group)
        for b in model.buffers():
            dist.broadcast(b.data, src=0, group=group)
        self.consolidate()                          # every replica starts from the owners' slices
        # peer-visible buffers of this rank: indices and gradient rows of its batch, per tower
        self.x_shared = [torch.zeros(self.B, h.n_tables, dtype=torch.int64, device=dev) for h in handles]
        self.dx_shared = [torch.zeros(self.B, h.n_tables * h.emb_dim, device=dev) for h in handles]
        named = {}
        for (t, k, c), r in zip(self.slices, self.owner):
            if r == self.rank:
                named[f"table{t}.{k}"] = handles[t].embeddings[k].weight.data
        for t in range(len(handles)):
            named[f"x{t}"], named[f"dx{t}"] = self.x_shared[t], self.dx_shared[t]
        self.peers = peers or CudaIpcPeers(group)
        self.views = self.peers.share(named)
        owner_of = dict(zip(self.slices, self.owner))
        for t, h in enumerate(handles):
            tables = [self.views[f"table{t}.{k}"][owner_of[(t, k, c)]]
                      for k in range(h.n_tables) for c in range(self.pieces[t])]
            h.row_source = self.kernels.make_row_source(h, tables, self.pieces[t], self.dx_shared[t])
        # owned slices grouped by tower (one embedding width per group); ONE sort covers all groups
        self.groups = []
        for t, h in enumerate(handles):
            width = h.emb_dim // self.pieces[t]
            owned = [dict(n_cols=h.n_tables, col=k, col0=c * width, grad=h.embeddings[k].weight.grad,
                          x_cat=self.views[f"x{t}"], dx_emb=self.views[f"dx{t}"])
                     for (tt, k, c), r in zip(self.slices, self.owner) if tt == t and r == self.rank]
            if owned:
                self.groups.append(dict(owned=owned, emb_dim=h.emb_dim, width=width))
        n_owned = sum(len(g["owned"]) for g in self.groups)
        self.scratch = self.kernels.make_scratch(n_owned, self.world, self.B, dev) if n_owned else None
        self._dirty = False          # owned gradient rows written by the last reduce, not yet re-zeroed
        self._presorted = False      # begin_step() already built and sorted this step's keys
        self._side = self.kernels.side_stream(dev) if dev.type == "cuda" else None
        self._token = torch.zeros(1, device=dev)
        dist.barrier(group=group)

    # ---- parameter views ----------------------------------------------------------------
    def owned_parameters(self) -> List[torch.nn.Parameter]:
        """Dense parameters plus the tables this rank owns a slice of: what this rank's optimiser should step
        (columns owned by another rank keep zero gradients here)."""
        mine = []
        for (t, k, c), r in zip(self.slices, self.owner):
            w = self.handles[t].embeddings[k].weight
            if r == self.rank and all(w is not m for m in mine):
                mine.append(w)
        return self.dense + mine

    def consolidate(self) -> None:
        """Broadcast every slice from its owner (start-up, checkpointing: ``state_dict()`` is then complete)."""
        for (t, k, c), r in zip(self.slices, self.owner):
            w = self.handles[t].embeddings[k].weight.data
            width = w.shape[1] // self.pieces[t]
            part = w[:, c * width:(c + 1) * width].contiguous()
            dist.broadcast(part, src=r, group=self.group)
            w[:, c * width:(c + 1) * width] = part

    # ---- per-step protocol --------------------------------------------------------------
    def begin_step(self, f_cat: torch.Tensor, c_cat: torch.Tensor) -> None:
        """Optional, before the forward: publish this step's indices right away, rendezvous once, and let every
        owner build and sort its (slice, index) keys on a side stream WHILE the step computes (the sort needs only
        indices; it is ~40 % of the owner-side reduce).  ``sync_gradients`` then only runs the segment reduce."""
        for t, x in enumerate((f_cat, c_cat)):
            self.x_shared[t].copy_(x)
        dist.all_reduce(self._token, group=self.group)      # every rank's indices are in place
        self.peers.fence()
        self.rezero()                                       # previous keys are consumed before the sort overwrites them
        if self.groups:
            if self._side is not None:
                main = torch.cuda.current_stream()
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self.kernels.peer_reduce(self.groups, self.world, self.B, self.scratch, phase=1)
            else:
                self.kernels.peer_reduce(self.groups, self.world, self.B, self.scratch, phase=1)
        self._presorted = True

    def sync_gradients(self, release: Optional[bool] = None) -> None:
        """After ``backward()``: all-reduce the tower gradients and let every owner reduce its slices' gradients.
        ``release``: run the closing rendezvous here.  Default: only when the step did not start with
        ``begin_step`` — with it, the next step's opening rendezvous already orders everything (every rank reaches it
        after its own reduce and optimiser step).  With an optimiser and no ``begin_step`` pass ``release=False`` and
        call ``release()`` after ``optimizer.step()``."""
        if release is None:
            release = not self._presorted
        for t, h in enumerate(self.handles):
            pg = h.table_grads
            if pg.pending is None:
                continue
            x_cat, dx_emb = pg.pending            # kept: a replayed CUDA graph refills the same buffers
            if not self._presorted:
                self.x_shared[t].copy_(x_cat)
            if dx_emb.data_ptr() != self.dx_shared[t].data_ptr():
                self.dx_shared[t].copy_(dx_emb)
        if self._presorted and self._side is not None:
            # joined BEFORE the all-reduce: entering it then implies this rank no longer reads the peers' indices,
            # so a peer that leaves the all-reduce may publish the next step's
            torch.cuda.current_stream().wait_stream(self._side)
        # one flat all-reduce for the tower parameters; it completes only after every rank has queued it behind its
        # backward, so it is also the barrier after which all peer buffers are complete
        allreduce_flat_([p.grad for p in self.dense] + [self._token], self.group)
        self.peers.fence()
        if self._presorted:
            if self.groups:
                self.kernels.peer_reduce(self.groups, self.world, self.B, self.scratch, phase=2)
                self._dirty = True
            self._presorted = False
        else:
            self.rezero()                         # rows of the previous step (its sorted keys are still in scratch)
            if self.groups:
                self.kernels.peer_reduce(self.groups, self.world, self.B, self.scratch, phase=0)
                self._dirty = True
        if release:
            self.release()

    def release(self) -> None:
        """Last rendezvous of the step: after it, every owner has consumed the peers' gradient rows (they may be
        overwritten) and, with an optimiser, every owner's updated table rows are visible to the next forward."""
        dist.all_reduce(self._token, group=self.group)

    def rezero(self) -> None:
        """Zero the owned gradient slices written by the last reduce.  Runs at the start of the next step's key
        build (not inside ``zero_grad_fast``, which a CUDA graph may have captured before the first reduce
        existed); call it directly to get clean ``.grad`` tables earlier."""
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if self.groups and (self._dirty or capturing):    # idempotent: a captured step always carries the re-zero node
            self.kernels.rezero(self.groups, self.world, self.B, self.scratch)
        self._dirty = False


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
        # ONE all-gather for both sides (the collectives are latency-bound at these sizes): [world, 2, n, Dp]
        both = gather_rows(torch.stack([fb, cb]).unsqueeze(0), group)
        fb_all = both[:, 0].reshape(world * n, -1)
        cb_all = both[:, 1].reshape(world * n, -1)
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
        world, n = dist.get_world_size(group), rs_f.shape[0]
        rs_both = gather_rows(torch.stack([rs_f, rs_c]).unsqueeze(0), group)    # one all-gather: [world, 2, n]
        rs_f_all = rs_both[:, 0].reshape(world * n)
        rs_c_all = rs_both[:, 1].reshape(world * n)
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
                       backend=None, merge: str = "fused") -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-``k`` columns (global indices) for this rank's rows when BOTH sides are row-sharded across ranks
    (BASELINE config 5).  The column shards are all-gathered (ragged shards allowed; 128 MB at 1M x 64 bf16) and

    * ``merge="fused"``  – the rank's rows are scored against the concatenated columns in ONE pass: the kernel's own
      candidate lists do the cross-shard merge, and the per-row selection / fp64 rescoring runs once;
    * ``merge="shards"`` – every shard is scored separately with its global column offset and the ``world`` partial
      lists go through the exact k-way merge (``cfm_topk_merge``, fp64 scores) — the form to use when the column
      side does not fit one GPU or arrives shard by shard.

    Both return the same lists (tests pin them against each other and the oracle)."""
    backend = backend or CudaScoringBackend()
    cols_all, counts = gather_ragged_rows(cols_local.contiguous(), group)
    if merge == "fused":
        s, i = backend.topk(rows_local, cols_all, k, scale, 0)
        return s.float(), i
    if merge != "shards":
        raise ValueError("merge must be 'fused' or 'shards'")
    parts_s, parts_i = [], []
    off = 0
    for c in counts:
        if c > 0:
            s, i = backend.topk(rows_local, cols_all[off:off + c], k, scale, off)
            parts_s.append(s)
            parts_i.append(i)
        off += c
    if len(parts_s) == 1:
        return parts_s[0].float(), parts_i[0]
    return backend.merge(torch.stack(parts_s), torch.stack(parts_i))
