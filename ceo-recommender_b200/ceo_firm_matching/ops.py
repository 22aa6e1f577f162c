"""Host side of the fused ops: autograd Functions over the libcfm_b200 C ABI.

torch is used for device memory, streams and autograd bookkeeping only; all
arithmetic of the hot path runs in the CUDA library (``_native.lib()``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _native as N


# ---------------------------------------------------------------------------------------
# small per-device state
# ---------------------------------------------------------------------------------------
class _DeviceState:
    """Per-device persistent buffers: error flag, reduction partials, dropout counter."""

    def __init__(self, device: torch.device):
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.partial = torch.zeros(4096, dtype=torch.float32, device=device)   # last-CTA-reduces scratch
        self.dropout_offset = 0


_states = {}


def _state(device: torch.device) -> _DeviceState:
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _states:
        _states[key] = _DeviceState(torch.device("cuda", key))
    return _states[key]


def raise_if_index_error(device: Optional[torch.device] = None) -> None:
    """Synchronising check of the device-side error word (mirrors torch's IndexError for a
    categorical index outside its embedding table — model.py:69,74)."""
    for key, st in list(_states.items()):
        if device is not None and device.index not in (None, key):
            continue
        flag = int(st.err_flag.item())
        if flag & 1:
            st.err_flag.zero_()
            raise IndexError("index out of range in self (categorical index outside its embedding table)")


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "ceo_firm_matching (B200 build) runs on CUDA only and has no CPU fallback; got a tensor on "
                f"{t.device}. Move the module and its inputs to 'cuda'.")


def next_dropout_offset(device: torch.device) -> Tuple[int, int]:
    """(seed, offset) of the next dropout draw: seed follows ``torch.manual_seed``, offset counts calls."""
    st = _state(device)
    st.dropout_offset += 1
    return torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, st.dropout_offset


# While a training step is being captured into a CUDA graph the per-call Python offset would be frozen into the
# graph; instead the kernels add a device-resident counter (advanced by one graph node per step) to the offset.
_graph_rng_counter: Optional[torch.Tensor] = None


def set_graph_rng_counter(counter: Optional[torch.Tensor]) -> None:
    global _graph_rng_counter
    _graph_rng_counter = counter


def advance_graph_rng_counter() -> None:
    if _graph_rng_counter is not None:
        with torch.cuda.device(_graph_rng_counter.device):
            N.check(N.lib().cfm_counter_advance(N.ptr(_graph_rng_counter), 1, N.stream_ptr()))


# ---------------------------------------------------------------------------------------
# towers
# ---------------------------------------------------------------------------------------
class TowerHandle:
    """Binds one tower's nn.Embedding list + nn.Sequential to the ``cfm_tower_t`` layout.

    The sub-modules stay ordinary torch modules (scripts call ``model.firm_tower(x)`` directly —
    deep_dive.py:80-91); the fused op just reads their parameters.
    """

    def __init__(self, embeddings: nn.ModuleList, seq: nn.Sequential, lin: Sequence[int],
                 bn: Sequence[Optional[int]], drop: Sequence[float], tower_id: int):
        self.embeddings, self.seq = embeddings, seq
        self.lin, self.bn, self.tower_id = tuple(lin), tuple(bn), tower_id
        self._drop = tuple(drop)
        self._scratch = None
        self._wimg = None
        self.use_tc = True                       # tcgen05 stage kernels where the layer shapes allow (False: mma.sync)
        self.precision = 0                       # 0: fp32-class (3xTF32), 1: single-pass TF32 tensor-core products
        self.table_grads: Optional["PersistentTableGrads"] = None
        self.row_source: Optional["StashedRows"] = None   # table-sharded mode: rows pulled into a local stash first

    # dims -------------------------------------------------------------------------------
    @property
    def n_tables(self) -> int:
        return len(self.embeddings)

    @property
    def emb_dim(self) -> int:
        return self.embeddings[0].embedding_dim if self.n_tables else 0

    def linear(self, i: int) -> nn.Linear:
        return self.seq[self.lin[i]]

    def batchnorm(self, i: int) -> Optional[nn.BatchNorm1d]:
        return self.seq[self.bn[i]] if self.bn[i] is not None else None

    def drop_p(self, i: int) -> float:
        """Dropout probability at site i, read live from the module so ``m.p = 0`` is honoured."""
        if self._drop[i] is None:
            return 0.0
        mod = self.seq[self._drop[i]]
        return float(mod.p)

    def params(self) -> List[torch.Tensor]:
        """Differentiable tensors in the canonical order the Function uses."""
        ps = [e.weight for e in self.embeddings]
        l1, l2, l3 = self.linear(0), self.linear(1), self.linear(2)
        b1, b2 = self.batchnorm(0), self.batchnorm(1)
        ps += [l1.weight, l1.bias, b1.weight, b1.bias, l2.weight, l2.bias]
        if b2 is not None:
            ps += [b2.weight, b2.bias]
        ps += [l3.weight, l3.bias]
        return ps

    def weight_images(self, device: torch.device, floats: int) -> torch.Tensor:
        """(hi, lo)-split swizzled weight images of the tcgen05 stage kernels; rebuilt by every forward call."""
        if self._wimg is None or self._wimg.numel() < floats or self._wimg.device != device:
            self._wimg = torch.empty(floats, dtype=torch.float32, device=device)
        return self._wimg

    def scratch(self, device: torch.device, floats: int) -> torch.Tensor:
        if self._scratch is None or self._scratch.numel() < floats or self._scratch.device != device:
            self._scratch = torch.empty(floats, dtype=torch.float32, device=device)
        return self._scratch


class _TowerCall:
    """Everything one fwd/bwd pair of a tower shares: the C struct and the tensors keeping it alive."""

    def __init__(self, h: TowerHandle, x_num: torch.Tensor, x_cat: torch.Tensor, B: int, training: bool = False):
        dev = x_num.device
        l1, l2, l3 = h.linear(0), h.linear(1), h.linear(2)
        b1, b2 = h.batchnorm(0), h.batchnorm(1)
        in_dim = l1.in_features
        n_num = in_dim - h.n_tables * h.emb_dim
        if x_num.shape != (B, n_num) or (h.n_tables and x_cat.shape != (B, h.n_tables)):
            raise RuntimeError(f"tower input shapes {tuple(x_num.shape)}, {tuple(x_cat.shape)} do not match "
                               f"[B,{n_num}] / [B,{h.n_tables}]")
        self.h = h
        self.B = B
        self.x_num, self.x_cat = x_num, x_cat
        self.h1_raw = torch.empty(B, l1.out_features, device=dev)
        self.h2_raw = torch.empty(B, l2.out_features, device=dev)
        self.out = torch.empty(B, l3.out_features, device=dev)
        self.bn1_stat = torch.empty(4, l1.out_features, device=dev)
        self.bn2_stat = torch.empty(4, l2.out_features, device=dev)
        t = N.Tower()
        t.n_num, t.n_tables, t.emb_dim = n_num, h.n_tables, h.emb_dim
        t.h1, t.h2, t.d_out = l1.out_features, l2.out_features, l3.out_features
        t.bn2 = 1 if b2 is not None else 0
        t.drop1, t.drop2 = h.drop_p(0), h.drop_p(1)
        t.tower_id = h.tower_id
        t.precision = h.precision
        t.x_num, t.x_cat = N.ptr(x_num), N.ptr(x_cat)
        if h.row_source is not None and h.n_tables:
            # rows come from (possibly NVLink peer-mapped) tables into a per-call stash [K, B, E]; the tower kernels
            # then see table k = stash[k] and index = row number, in forward and in the stage-1 backward
            self.stash, self.stash_index = h.row_source.gather(x_cat, B)
            t.x_cat = N.ptr(self.stash_index)
            for i in range(h.n_tables):
                t.tables[i] = N.ptr(self.stash[i])
                t.table_rows[i] = B
        else:
            for i, e in enumerate(h.embeddings):
                t.tables[i] = N.ptr(e.weight)
                t.table_rows[i] = e.num_embeddings
        t.w1, t.b1, t.w2, t.b2, t.w3, t.b3 = (N.ptr(x) for x in (l1.weight, l1.bias, l2.weight, l2.bias,
                                                                  l3.weight, l3.bias))
        t.bn1_w, t.bn1_b = N.ptr(b1.weight), N.ptr(b1.bias)
        t.bn1_rm, t.bn1_rv, t.bn1_nbt = N.ptr(b1.running_mean), N.ptr(b1.running_var), N.ptr(b1.num_batches_tracked)
        if b2 is not None:
            t.bn2_w, t.bn2_b = N.ptr(b2.weight), N.ptr(b2.bias)
            t.bn2_rm, t.bn2_rv, t.bn2_nbt = (N.ptr(b2.running_mean), N.ptr(b2.running_var),
                                             N.ptr(b2.num_batches_tracked))
        t.h1_raw, t.h2_raw, t.out = N.ptr(self.h1_raw), N.ptr(self.h2_raw), N.ptr(self.out)
        t.bn1_stat, t.bn2_stat = N.ptr(self.bn1_stat), N.ptr(self.bn2_stat)
        per_cta = N.lib().cfm_tower_scratch_floats(C.byref(t))
        t.scratch = N.ptr(h.scratch(dev, per_cta * N.device_info()["tower_ctas"]))
        if h.use_tc:
            t.wimg = N.ptr(h.weight_images(dev, N.lib().cfm_tower_wimg_floats(C.byref(t))))
            self.a1 = torch.empty(B, l1.out_features, device=dev)
            self.a2 = torch.empty(B, l2.out_features, device=dev)
            t.a1, t.a2 = N.ptr(self.a1), N.ptr(self.a2)
            if training and h.n_tables:
                # gathered stage-1 inputs, kept for the backward (sequential re-read instead of a second random gather)
                self.xstash = torch.empty(N.lib().cfm_tower_xstash_floats(C.byref(t), B), device=dev)
                t.xstash = N.ptr(self.xstash)
        self.struct = t


def _as_index(x_cat: torch.Tensor) -> torch.Tensor:
    # explain.py:53,59 hands categoricals over as float -> .long(); accept any integer/float dtype
    return x_cat if x_cat.dtype == torch.int64 else x_cat.long()


class TowersFunction(torch.autograd.Function):
    """Forward/backward of one or two towers over the same batch (model.py:69-76; structural_model.py:120-127).

    apply(handles, training, seed, offset, x_num_0, x_cat_0[, x_num_1, x_cat_1], *params) -> out_0[, out_1]
    """

    @staticmethod
    def forward(ctx, handles, training, seed, offset, *tensors):
        n = len(handles)
        xs = tensors[:2 * n]
        _require_cuda(*tensors)
        B = xs[0].shape[0]
        dev = xs[0].device
        calls = []
        with torch.cuda.device(dev):
            for i, h in enumerate(handles):
                x_num = xs[2 * i].contiguous().float()
                x_cat = _as_index(xs[2 * i + 1]).contiguous()
                calls.append(_TowerCall(h, x_num, x_cat, B, training))
            arr = (N.Tower * n)(*[c.struct for c in calls])
            rng_dev = _graph_rng_counter if training else None
            joint = _joint_for(calls) if (training and any(ctx.needs_input_grad)) else None
            if joint is not None and all(c.h.row_source is None for c in calls):
                joint.presort([c.x_cat for c in calls if c.h.n_tables], B)
            N.check(N.lib().cfm_towers_fwd(arr, n, B, 1 if training else 0, seed, offset, N.ptr(rng_dev),
                                           N.ptr(_state(dev).err_flag), N.stream_ptr()))
        ctx.calls, ctx.arr, ctx.training, ctx.seed, ctx.offset = calls, arr, training, seed, offset
        ctx.rng_dev = rng_dev
        ctx.handles = handles
        ctx.needs_xnum = [xs[2 * i].requires_grad for i in range(n)]
        outs = tuple(c.out for c in calls)
        return outs if n > 1 else outs[0]

    @staticmethod
    def backward(ctx, *g_outs):
        calls, n = ctx.calls, len(ctx.calls)
        dev = calls[0].out.device
        B = calls[0].B
        per_tower = []
        grads_arr = (N.TowerGrads * n)()
        with torch.cuda.device(dev):
            for i, c in enumerate(calls):
                h, t = c.h, c.struct
                g = g_outs[i]
                g = torch.zeros_like(c.out) if g is None else g.contiguous().float()
                K_in = t.n_num + t.n_tables * t.emb_dim
                d = dict(
                    g_out=g,
                    dw1=torch.empty(t.h1, K_in, device=dev), db1=torch.empty(t.h1, device=dev),
                    dw2=torch.empty(t.h2, t.h1, device=dev), db2=torch.empty(t.h2, device=dev),
                    dw3=torch.empty(t.d_out, t.h2, device=dev), db3=torch.empty(t.d_out, device=dev),
                    dbn1_w=torch.empty(t.h1, device=dev), dbn1_b=torch.empty(t.h1, device=dev),
                    dbn2_w=torch.empty(t.h2, device=dev) if t.bn2 else None,
                    dbn2_b=torch.empty(t.h2, device=dev) if t.bn2 else None,
                    dy1=torch.empty(B, t.h1, device=dev), dy2=torch.empty(B, t.h2, device=dev),
                    dx_emb=_dx_emb_buffer(h, B, t.n_tables * t.emb_dim, dev) if t.n_tables else None,
                    dx_num=torch.empty(B, t.n_num, device=dev) if ctx.needs_xnum[i] else None,
                )
                for k, v in d.items():
                    setattr(grads_arr[i], k, N.ptr(v))
                per_tower.append(d)
            N.check(N.lib().cfm_towers_bwd(ctx.arr, grads_arr, n, B, 1 if ctx.training else 0, ctx.seed, ctx.offset,
                                           N.ptr(ctx.rng_dev), N.stream_ptr()))
            out: List[Optional[torch.Tensor]] = []
            for i, (c, d) in enumerate(zip(calls, per_tower)):
                out += [d["dx_num"], None]
            joint = _joint_for(calls)
            if joint is not None:                     # one sort for all towers, straight into the persistent grads
                joint.reduce([(c.x_cat, d["dx_emb"]) for c, d in zip(calls, per_tower)], B)
            for c, d in zip(calls, per_tower):
                h, t = c.h, c.struct
                if joint is not None:
                    table_grads = [None] * t.n_tables
                else:
                    table_grads = embedding_grads(h, c.x_cat, d["dx_emb"], B) if t.n_tables else []
                out += table_grads
                out += [d["dw1"], d["db1"], d["dbn1_w"], d["dbn1_b"], d["dw2"], d["db2"]]
                if t.bn2:
                    out += [d["dbn2_w"], d["dbn2_b"]]
                out += [d["dw3"], d["db3"]]
        return (None, None, None, None, *out)


def run_towers(handles: Sequence[TowerHandle], inputs: Sequence[Tuple[torch.Tensor, torch.Tensor]], training: bool):
    """Convenience wrapper: gathers parameters, draws the dropout (seed, offset) and applies the Function."""
    dev = inputs[0][0].device
    _require_cuda(*(t for pair in inputs for t in pair))
    seed, offset = next_dropout_offset(dev) if training else (0, 0)
    flat = [t for pair in inputs for t in pair]
    params = [p for h in handles for p in h.params()]
    _require_cuda(*params)
    return TowersFunction.apply(list(handles), training, seed, offset, *flat, *params)


# ---------------------------------------------------------------------------------------
# table-sharded mode: embedding rows pulled from local / NVLink peer-mapped tables into a local stash
# ---------------------------------------------------------------------------------------
class StashedRows:
    """Row source of a tower whose tables may live on other GPUs of the node (``distributed.TableShardedTwoTower``).

    ``tables[k * pieces + c]`` is the full ``[rows_k, E]`` weight of table ``k`` as held by the owner of its column
    slice ``c``: the local parameter, or a CUDA-IPC mapping of the owner's parameter.  ``gather`` copies the rows a batch needs into a fresh ``[K, B, E]`` stash, so a training step
    crosses NVLink once for the forward AND the stage-1 backward (which rebuilds its input tile from the same stash).
    ``dx_emb`` (optional) is a fixed ``[B, K*E]`` buffer the backward writes the per-pair gradient rows into; the
    table owners read it in place through their own peer mappings."""

    def __init__(self, handle: TowerHandle, tables: Sequence[torch.Tensor], pieces: int = 1,
                 dx_emb: Optional[torch.Tensor] = None):
        self.h = handle
        self.tables, self.pieces = list(tables), int(pieces)
        if len(self.tables) != handle.n_tables * self.pieces:
            raise ValueError("one (local or peer-mapped) tensor per (table, column slice) is required")
        self.dx_emb = dx_emb
        self._index = {}

    def index_for(self, B: int, device: torch.device) -> torch.Tensor:
        if B not in self._index:
            self._index[B] = torch.arange(B, dtype=torch.int64, device=device).unsqueeze(1).repeat(
                1, self.h.n_tables).contiguous()
        return self._index[B]

    def gather(self, x_cat: torch.Tensor, B: int) -> Tuple[torch.Tensor, torch.Tensor]:
        h, dev = self.h, x_cat.device
        stash = torch.empty(h.n_tables, B, h.emb_dim, device=dev)
        ptrs = (C.c_void_p * len(self.tables))(*[N.ptr(t) for t in self.tables])
        rows = (N.i64 * h.n_tables)(*[e.num_embeddings for e in h.embeddings])
        N.check(N.lib().cfm_emb_gather_rows(N.ptr(x_cat), B, h.n_tables, h.emb_dim, self.pieces, ptrs, rows,
                                            N.ptr(stash), N.ptr(_state(dev).err_flag), N.stream_ptr()))
        return stash, self.index_for(B, dev)


def _dx_emb_buffer(h: TowerHandle, B: int, width: int, dev: torch.device) -> torch.Tensor:
    rs = h.row_source
    if rs is not None and rs.dx_emb is not None:
        if tuple(rs.dx_emb.shape) != (B, width):
            raise RuntimeError(f"table-sharded mode was set up for batches of {rs.dx_emb.shape[0]} rows per rank, "
                               f"got {B}")
        return rs.dx_emb
    return torch.empty(B, width, device=dev)


# ---------------------------------------------------------------------------------------
# embedding gradients: sorted-segment reduce into dense [n_i, E] tables
# ---------------------------------------------------------------------------------------
class _SortScratch:
    def __init__(self, n_items: int, n_tables: int, B: int, device: torch.device):
        self.n_items = n_items
        self.keys_tmp = torch.empty(n_items, dtype=torch.int64, device=device)
        self.vals_tmp = torch.empty(n_items, dtype=torch.int32, device=device)
        self.keys_sorted = torch.empty(n_items, dtype=torch.int64, device=device)
        self.vals_sorted = torch.empty(n_items, dtype=torch.int32, device=device)
        self.tmp_bytes = N.lib().cfm_emb_grad_tmp_bytes(n_tables, B)
        self.tmp = torch.empty(self.tmp_bytes, dtype=torch.uint8, device=device)


class PersistentTableGrads:
    """Keeps each table's dense ``.grad`` allocated across steps and re-zeroes only the rows the previous
    step touched (their sorted keys are kept), so the result is the exact dense gradient torch produces
    (``nn.Embedding(sparse=False)``, model.py:24-33) without a full-table memset every step.
    ``rezero()`` reads the sorted keys of the last backward; it runs either at the start of the next step
    (``zero_grad_fast``) or right after the optimiser step — both orders are CUDA-graph capturable because
    the keys are consumed before the next sort overwrites them."""

    def __init__(self, handle: TowerHandle):
        self.h = handle
        self.scratches = {}                      # n_items -> _SortScratch (kept alive: graphs hold raw pointers)
        self.prev: Optional[_SortScratch] = None
        self.defer = False                       # data parallelism: keep (indices, rows) for the cross-rank reduce
        self.pending = None
        self.rezero_hook = None                  # table-sharded mode: the owner re-zeroes its tables itself
        self.joint: Optional["JointTableGrads"] = None   # set when all towers of a model reduce in one sort
        for e in handle.embeddings:
            e.weight.grad = torch.zeros_like(e.weight)

    def scratch_for(self, n_items: int, B: int, device: torch.device) -> "_SortScratch":
        if n_items not in self.scratches:
            self.scratches[n_items] = _SortScratch(n_items, self.h.n_tables, B, device)
        return self.scratches[n_items]

    def rezero(self) -> None:
        if self.rezero_hook is not None:
            self.rezero_hook()
            return
        if self.joint is not None:
            self.joint.rezero()                  # idempotent: the first tower's call does the work
        if self.prev is None:
            return
        h = self.h
        ptrs = (C.c_void_p * h.n_tables)(*[N.ptr(e.weight.grad) for e in h.embeddings])
        rows = (N.i64 * h.n_tables)(*[e.num_embeddings for e in h.embeddings])
        N.check(N.lib().cfm_emb_grad_rezero(ptrs, rows, h.n_tables, h.emb_dim, N.ptr(self.prev.keys_sorted),
                                            self.prev.n_items, N.stream_ptr()))
        self.prev = None


class JointTableGrads:
    """All towers of one model reduce their embedding gradients with ONE radix sort per step
    (``cfm_emb_grad_joint_reduce``): at these sizes the sort passes are latency-bound, so one sort over both towers'
    (table, index) pairs costs about half of two.  Results are bitwise those of the per-tower reduce."""

    def __init__(self, handles: Sequence[TowerHandle]):
        self.handles = [h for h in handles if h.n_tables]
        self.total_tables = sum(h.n_tables for h in self.handles)
        self.scratches = {}
        self.prev = None                          # (scratch, B) of the last reduce, until re-zeroed
        self.presorted = None                     # (index tensor addresses, B) whose sorted keys are in the scratch
        self._side: Optional[torch.cuda.Stream] = None
        for h in self.handles:
            h.table_grads.joint = self

    def usable(self) -> bool:
        return (self.total_tables <= N.CFM_MAX_TABLES and
                all(h.table_grads is not None and h.table_grads.joint is self and not h.table_grads.defer and
                    h.table_grads.rezero_hook is None for h in self.handles))

    def _groups(self, inputs):
        arr = (N.EmbGroup * len(self.handles))()
        for g, h in enumerate(self.handles):
            if inputs is not None:
                arr[g].x_cat, arr[g].dx_emb = N.ptr(inputs[g][0]), N.ptr(inputs[g][1])
            arr[g].n_tables, arr[g].emb_dim = h.n_tables, h.emb_dim
            for i, e in enumerate(h.embeddings):
                arr[g].grad_tables[i] = N.ptr(e.weight.grad)
                arr[g].table_rows[i] = e.num_embeddings
        return arr

    def _scratch(self, B: int, dev: torch.device) -> "_SortScratch":
        if B not in self.scratches:
            self.scratches[B] = _SortScratch(B * self.total_tables, self.total_tables, B, dev)
        return self.scratches[B]

    def _call(self, inputs, B: int, phase: int, sc: "_SortScratch") -> None:
        N.check(N.lib().cfm_emb_grad_joint_reduce(self._groups(inputs), len(self.handles), B, phase, N.ptr(sc.keys_tmp),
                                                  N.ptr(sc.vals_tmp), N.ptr(sc.keys_sorted), N.ptr(sc.vals_sorted),
                                                  N.ptr(sc.tmp), sc.tmp_bytes, N.stream_ptr()))

    def presort(self, x_cats: Sequence[torch.Tensor], B: int) -> None:
        """Key build + radix sort of this step's (table, index) pairs on a side stream, started by the forward: the
        sort needs the indices only, so it runs beside the towers instead of after the backward (under graph capture
        the fork and the join become graph edges).  Skipped while the rows of an earlier backward still await their
        re-zeroing: that needs the previous sorted keys."""
        if self.prev is not None:
            return
        dev = x_cats[0].device
        sc = self._scratch(B, dev)
        main = torch.cuda.current_stream(dev)
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(dev)
        side = self._side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._call([(x, None) for x in x_cats], B, 1, sc)
        for x in x_cats:
            x.record_stream(side)
        self.presorted = (tuple(x.data_ptr() for x in x_cats), B)

    def reduce(self, inputs, B: int) -> None:
        if self.prev is not None:
            raise RuntimeError("persistent table grads: backward ran twice without zero_grad_fast() in between")
        dev = inputs[0][1].device
        sc = self._scratch(B, dev)
        if self.presorted == (tuple(x.data_ptr() for x, _ in inputs), B):
            torch.cuda.current_stream(dev).wait_stream(self._side)
            self._call(inputs, B, 2, sc)
        else:
            if self._side is not None:               # a sort for other inputs may still be running on the scratch
                torch.cuda.current_stream(dev).wait_stream(self._side)
            self._call(inputs, B, 0, sc)
        self.presorted = None
        self.prev = (sc, B)

    def rezero(self) -> None:
        if self.prev is None:
            return
        sc, B = self.prev
        N.check(N.lib().cfm_emb_grad_joint_rezero(self._groups(None), len(self.handles), B, N.ptr(sc.keys_sorted),
                                                  N.stream_ptr()))
        self.prev = None


def _joint_for(calls) -> Optional[JointTableGrads]:
    """The joint reducer when every tower of this call belongs to it and nothing defers the reduce."""
    pgs = [c.h.table_grads for c in calls if c.h.n_tables]
    if len(pgs) < 2 or any(pg is None or pg.joint is None for pg in pgs):
        return None
    joint = pgs[0].joint
    if any(pg.joint is not joint for pg in pgs) or len(joint.handles) != len(pgs) or not joint.usable():
        return None
    if any(a is not b for a, b in zip(joint.handles, [c.h for c in calls if c.h.n_tables])):
        return None
    return joint


def _segment_reduce(h: TowerHandle, x_cat: torch.Tensor, dx_emb: torch.Tensor, B: int,
                    targets: Sequence[torch.Tensor], scratch: _SortScratch) -> None:
    ptrs = (C.c_void_p * h.n_tables)(*[N.ptr(t) for t in targets])
    rows = (N.i64 * h.n_tables)(*[e.num_embeddings for e in h.embeddings])
    N.check(N.lib().cfm_emb_grad_segment_reduce(
        N.ptr(x_cat), N.ptr(dx_emb), B, h.n_tables, h.emb_dim, ptrs, rows, N.ptr(scratch.keys_tmp),
        N.ptr(scratch.vals_tmp), N.ptr(scratch.keys_sorted), N.ptr(scratch.vals_sorted),
        N.ptr(scratch.tmp), scratch.tmp_bytes, N.stream_ptr()))


def embedding_grads(h: TowerHandle, x_cat: torch.Tensor, dx_emb: torch.Tensor, B: int) -> List[Optional[torch.Tensor]]:
    """Dense gradient of every table of tower ``h`` (aten::embedding_dense_backward semantics)."""
    n_items = B * h.n_tables
    pg = h.table_grads
    if pg is not None:
        # fast path: write into the persistent .grad buffers (rows touched last step were re-zeroed by
        # PersistentTableGrads.rezero(), called from zero_grad), autograd gets no tensor for the tables
        if pg.prev is not None:
            raise RuntimeError("persistent table grads: backward ran twice without zero_grad_fast() in between")
        if pg.defer:                                  # reduced later over the all-gathered global batch
            pg.pending = (x_cat, dx_emb)
            return [None] * h.n_tables
        scratch = pg.scratch_for(n_items, B, dx_emb.device)
        _segment_reduce(h, x_cat, dx_emb, B, [e.weight.grad for e in h.embeddings], scratch)
        pg.prev = scratch
        return [None] * h.n_tables
    scratch = _SortScratch(n_items, h.n_tables, B, dx_emb.device)
    grads = [torch.zeros_like(e.weight) for e in h.embeddings]
    _segment_reduce(h, x_cat, dx_emb, B, grads, scratch)
    return grads


def reduce_table_grads(h: TowerHandle, x_cat: torch.Tensor, dx_emb: torch.Tensor) -> None:
    """Segment-reduce externally supplied (index, gradient-row) pairs — e.g. the all-gathered global batch of a
    data-parallel step — into the tower's persistent dense table gradients."""
    pg = h.table_grads
    B = x_cat.shape[0]
    if pg.prev is not None:
        # rows of the previous reduce; done here and not only in zero_grad_fast because a CUDA graph captured
        # before the first reduce existed has no re-zero node to replay
        pg.rezero()
    scratch = pg.scratch_for(B * h.n_tables, B, dx_emb.device)
    _segment_reduce(h, x_cat.contiguous(), dx_emb.contiguous(), B, [e.weight.grad for e in h.embeddings], scratch)
    pg.prev = scratch


# ---------------------------------------------------------------------------------------
# cosine head (model.py:79-87 / contrastive.py:64-70,92-93) with optional fused weighted MSE (training.py:52)
# ---------------------------------------------------------------------------------------
class CosineHeadFunction(torch.autograd.Function):
    """apply(u, v, logit_scale, eps, want_unit) -> score [B,1] (+ u_hat, v_hat when want_unit)."""

    @staticmethod
    def forward(ctx, u, v, logit_scale, eps, want_unit):
        _require_cuda(u, v, logit_scale)
        u, v = u.contiguous(), v.contiguous()
        B, D = u.shape
        score = torch.empty(B, 1, device=u.device)
        u_hat = torch.empty_like(u) if want_unit else None
        v_hat = torch.empty_like(v) if want_unit else None
        with torch.cuda.device(u.device):
            N.check(N.lib().cfm_cosine_head_fwd(N.ptr(u), N.ptr(v), N.ptr(logit_scale), B, D, eps, N.ptr(score),
                                                N.ptr(u_hat), N.ptr(v_hat), None, None, None, None, N.stream_ptr()))
        ctx.save_for_backward(u, v, logit_scale)
        ctx.eps = eps
        return (score, u_hat, v_hat) if want_unit else score

    @staticmethod
    def backward(ctx, d_score, d_uhat=None, d_vhat=None):
        u, v, logit_scale = ctx.saved_tensors
        B, D = u.shape
        du, dv = torch.empty_like(u), torch.empty_like(v)
        dls = torch.empty((), device=u.device)            # assigned (not accumulated) by the kernel's last-CTA finish
        d_score = d_score.contiguous() if d_score is not None else None
        d_uhat = d_uhat.contiguous() if d_uhat is not None else None
        d_vhat = d_vhat.contiguous() if d_vhat is not None else None
        with torch.cuda.device(u.device):
            N.check(N.lib().cfm_cosine_head_bwd(N.ptr(u), N.ptr(v), N.ptr(logit_scale), N.ptr(d_score), N.ptr(d_uhat),
                                                N.ptr(d_vhat), None, None, None, B, D, ctx.eps, N.ptr(du), N.ptr(dv),
                                                N.ptr(dls), N.ptr(_state(u.device).partial), N.stream_ptr()))
        return du, dv, dls, None, None


class CosineMSEFunction(torch.autograd.Function):
    """apply(u, v, logit_scale, target, weights, eps) -> (loss, score): loss = mean(w (s - t)^2)."""

    @staticmethod
    def forward(ctx, u, v, logit_scale, target, weights, eps):
        _require_cuda(u, v, logit_scale, target, weights)
        u, v = u.contiguous(), v.contiguous()
        target, weights = target.contiguous().float(), weights.contiguous().float()
        B, D = u.shape
        score = torch.empty(B, 1, device=u.device)
        loss = torch.empty((), device=u.device)           # assigned by the kernel's last-CTA finish
        with torch.cuda.device(u.device):
            N.check(N.lib().cfm_cosine_head_fwd(N.ptr(u), N.ptr(v), N.ptr(logit_scale), B, D, eps, N.ptr(score), None,
                                                None, N.ptr(target), N.ptr(weights), N.ptr(loss),
                                                N.ptr(_state(u.device).partial), N.stream_ptr()))
        ctx.save_for_backward(u, v, logit_scale, target, weights)
        ctx.eps = eps
        ctx.mark_non_differentiable(score)
        return loss, score

    @staticmethod
    def backward(ctx, g_loss, _g_score):
        u, v, logit_scale, target, weights = ctx.saved_tensors
        B, D = u.shape
        du, dv = torch.empty_like(u), torch.empty_like(v)
        dls = torch.empty((), device=u.device)
        g_loss = g_loss.contiguous().float()
        with torch.cuda.device(u.device):
            N.check(N.lib().cfm_cosine_head_bwd(N.ptr(u), N.ptr(v), N.ptr(logit_scale), None, None, None,
                                                N.ptr(target), N.ptr(weights), N.ptr(g_loss), B, D, ctx.eps,
                                                N.ptr(du), N.ptr(dv), N.ptr(dls), N.ptr(_state(u.device).partial),
                                                N.stream_ptr()))
        return du, dv, dls, None, None, None


# ---------------------------------------------------------------------------------------
# projection heads of the contrastive model (contrastive.py:41-50, 96-97)
# ---------------------------------------------------------------------------------------
class ProjectorHeadsFunction(torch.autograd.Function):
    """apply(eps, x_0, w1_0, b1_0, w2_0, b2_0, x_1, ...) -> (out_0, out_1, ...):
    out_i = normalize(W2_i relu(W1_i x_i + b1_i) + b2_i), all heads over the same B rows in one launch."""

    @staticmethod
    def forward(ctx, eps, *tensors):
        _require_cuda(*tensors)
        assert len(tensors) % 5 == 0 and tensors
        n = len(tensors) // 5
        tensors = [t.contiguous().float() for t in tensors]
        B, dev = tensors[0].shape[0], tensors[0].device
        heads = (N.Projector * n)()
        outs, saved = [], []
        for i in range(n):
            x, w1, b1, w2, b2 = tensors[5 * i:5 * i + 5]
            if x.shape != (B, w1.shape[1]) or w2.shape[1] != w1.shape[0] or b1.shape != (w1.shape[0],) or b2.shape != (w2.shape[0],):
                raise ValueError("projector head %d: inconsistent shapes" % i)
            hid = torch.empty(B, w1.shape[0], device=dev)
            raw = torch.empty(B, w2.shape[0], device=dev)
            out = torch.empty(B, w2.shape[0], device=dev)
            h = heads[i]
            h.d_in, h.d_hid, h.d_out = w1.shape[1], w1.shape[0], w2.shape[0]
            h.w1, h.b1, h.w2, h.b2 = N.ptr(w1), N.ptr(b1), N.ptr(w2), N.ptr(b2)
            h.x, h.hid, h.raw, h.out = N.ptr(x), N.ptr(hid), N.ptr(raw), N.ptr(out)
            outs.append(out)
            saved += [x, w1, b1, w2, b2, hid, raw]
        with torch.cuda.device(dev):
            N.check(N.lib().cfm_projector_fwd(heads, n, B, float(eps), N.stream_ptr()))
        ctx.save_for_backward(*saved)
        ctx.eps, ctx.n = float(eps), n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *g_outs):
        n, saved = ctx.n, ctx.saved_tensors
        B, dev = saved[0].shape[0], saved[0].device
        heads, grads = (N.Projector * n)(), (N.ProjectorGrads * n)()
        result, keep = [None], []
        for i in range(n):
            x, w1, b1, w2, b2, hid, raw = saved[7 * i:7 * i + 7]
            g = g_outs[i]
            g = torch.zeros_like(raw) if g is None else g.contiguous().float()
            h, gr = heads[i], grads[i]
            h.d_in, h.d_hid, h.d_out = w1.shape[1], w1.shape[0], w2.shape[0]
            h.w1, h.b1, h.w2, h.b2 = N.ptr(w1), N.ptr(b1), N.ptr(w2), N.ptr(b2)
            h.x, h.hid, h.raw, h.out = N.ptr(x), N.ptr(hid), N.ptr(raw), None
            dx = torch.empty_like(x) if ctx.needs_input_grad[1 + 5 * i] else None
            dw1, db1, dw2, db2 = (torch.empty_like(t) for t in (w1, b1, w2, b2))
            scratch = torch.empty(int(N.lib().cfm_projector_scratch_floats(C.byref(h), B)), device=dev)
            gr.g_out, gr.dx, gr.dw1, gr.db1, gr.dw2, gr.db2 = N.ptr(g), N.ptr(dx), N.ptr(dw1), N.ptr(db1), N.ptr(dw2), N.ptr(db2)
            gr.scratch = N.ptr(scratch)
            keep += [g, scratch]
            result += [dx, dw1, db1, dw2, db2]
        with torch.cuda.device(dev):
            N.check(N.lib().cfm_projector_bwd(heads, grads, n, B, ctx.eps, N.stream_ptr()))
        return tuple(result)


def projector_heads(pairs, eps: float = 1e-12):
    """``pairs``: [(x, nn.Sequential(Linear, ReLU, Linear)), ...] -> tuple of L2-normalised projections
    (contrastive.py:41-50 applied at :88-97).  The sub-modules only hand over their parameters."""
    flat = []
    for x, seq in pairs:
        lin1, lin2 = seq[0], seq[2]
        flat += [x, lin1.weight, lin1.bias, lin2.weight, lin2.bias]
    return ProjectorHeadsFunction.apply(eps, *flat)


# ---------------------------------------------------------------------------------------
# structural head (structural_model.py:130-141, structural_training.py:75-77)
# ---------------------------------------------------------------------------------------
class StructuralHeadFunction(torch.autograd.Function):
    """apply(c_logits, f_logits, A) -> expected_match [B,1]."""

    @staticmethod
    def forward(ctx, c_logits, f_logits, A):
        _require_cuda(c_logits, f_logits, A)
        c_logits, f_logits = c_logits.contiguous(), f_logits.contiguous()
        B = c_logits.shape[0]
        match = torch.empty(B, 1, device=c_logits.device)
        with torch.cuda.device(c_logits.device):
            N.check(N.lib().cfm_structural_head(N.ptr(c_logits), N.ptr(f_logits), N.ptr(A), None, None, None, B, 0.0,
                                                N.ptr(match), None, None, None,
                                                N.ptr(_state(c_logits.device).partial), N.stream_ptr()))
        ctx.save_for_backward(c_logits, f_logits, A)
        return match

    @staticmethod
    def backward(ctx, d_match):
        c_logits, f_logits, A = ctx.saved_tensors
        B = c_logits.shape[0]
        dc, df = torch.empty_like(c_logits), torch.empty_like(f_logits)
        scratch_match = torch.empty(B, device=c_logits.device)
        d_match = d_match.contiguous().float()
        with torch.cuda.device(c_logits.device):
            N.check(N.lib().cfm_structural_head(N.ptr(c_logits), N.ptr(f_logits), N.ptr(A), None, None,
                                                N.ptr(d_match), B, 0.0, N.ptr(scratch_match), None, N.ptr(dc),
                                                N.ptr(df), N.ptr(_state(c_logits.device).partial), N.stream_ptr()))
        return dc, df, None


class StructuralKLFunction(torch.autograd.Function):
    """apply(c_logits, f_logits, A, target_ceo, target_firm) -> loss: KL(batchmean) of both sides, summed.
    The gradients w.r.t. both logit sets are produced by the same kernel launch as the loss."""

    @staticmethod
    def forward(ctx, c_logits, f_logits, A, target_ceo, target_firm):
        _require_cuda(c_logits, f_logits, A, target_ceo, target_firm)
        c_logits, f_logits = c_logits.contiguous(), f_logits.contiguous()
        target_ceo, target_firm = target_ceo.contiguous().float(), target_firm.contiguous().float()
        B = c_logits.shape[0]
        dev = c_logits.device
        loss = torch.zeros((), device=dev)
        match = torch.empty(B, device=dev)
        need_grad = c_logits.requires_grad or f_logits.requires_grad
        dc = torch.empty_like(c_logits) if need_grad else None
        df = torch.empty_like(f_logits) if need_grad else None
        with torch.cuda.device(dev):
            N.check(N.lib().cfm_structural_head(N.ptr(c_logits), N.ptr(f_logits), N.ptr(A), N.ptr(target_ceo),
                                                N.ptr(target_firm), None, B, 1.0, N.ptr(match), N.ptr(loss),
                                                N.ptr(dc), N.ptr(df), N.ptr(_state(dev).partial), N.stream_ptr()))
        ctx.dc, ctx.df = dc, df
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        # d loss / d logits was computed with kl_scale = 1; chain with the incoming scalar
        if ctx.dc is None:
            return None, None, None, None, None
        return ctx.dc * g_loss, ctx.df * g_loss, None, None, None


def dropout_mask(B: int, width: int, p: float, tower_id: int, site: int, seed: int, offset: int,
                 device: torch.device) -> torch.Tensor:
    """Materialise the keep-mask the tower kernels regenerate from (seed, offset) — differential tests only."""
    mask = torch.empty(B, width, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        N.check(N.lib().cfm_dropout_mask(N.ptr(mask), B, width, p, tower_id, site, seed, offset, N.stream_ptr()))
    return mask


# ---------------------------------------------------------------------------------------
# InfoNCE on the tensor cores (contrastive.py:102-138)
# ---------------------------------------------------------------------------------------
def padded_width(d: int) -> int:
    dp = 64 * ((d + 63) // 64)
    if dp > 128:
        raise N.CfmError(f"feature width {d} not supported by the tcgen05 similarity kernels (max 128)")
    return dp


def pack_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32/bf16 [R, D] -> bf16 [R, Dp] with zero padding to a multiple of 64 features (TMA/UMMA granularity)."""
    _require_cuda(x)
    R, D = x.shape
    dp = padded_width(D)
    if x.dtype == torch.bfloat16 and dp == D and x.is_contiguous():
        return x
    out = torch.empty(R, dp, dtype=torch.bfloat16, device=x.device)
    xf = x.detach().float().contiguous()
    with torch.cuda.device(x.device):
        N.check(N.lib().cfm_pack_rows_bf16(N.ptr(xf), R, D, dp, N.ptr(out), N.stream_ptr()))
    return out


def pack_f16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [R, D] -> fp16 [R, Dp], zero padded (operands of the all-pairs filter pass when they fit fp16's range)."""
    _require_cuda(x)
    R, D = x.shape
    dp = padded_width(D)
    out = torch.empty(R, dp, dtype=torch.float16, device=x.device)
    xf = x.detach().float().contiguous()
    with torch.cuda.device(x.device):
        N.check(N.lib().cfm_pack_rows_f16(N.ptr(xf), R, D, dp, N.ptr(out), N.stream_ptr()))
    return out


def simtile_scores(xb: torch.Tensor, yb: torch.Tensor) -> torch.Tensor:
    """Raw tensor-core score tile X . Y^T as fp32 [R, C] (parity/debug aid)."""
    out = torch.empty(xb.shape[0], yb.shape[0], device=xb.device)
    with torch.cuda.device(xb.device):
        N.check(N.lib().cfm_simtile_scores(N.ptr(xb), N.ptr(yb), xb.shape[0], yb.shape[0], xb.shape[1], N.ptr(out),
                                           N.stream_ptr()))
    return out


def infonce_rowsum(xb: torch.Tensor, yb: torch.Tensor, temperature: float, diag_offset: int = 0, want_diag: bool = True):
    """rowsum[i] = sum_j exp((x_i.y_j - 1)/T) and diag[i] = x_i . y_(i+diag_offset), S never materialised."""
    R, C, dp = xb.shape[0], yb.shape[0], xb.shape[1]
    dev = xb.device
    chunks = N.lib().cfm_simtile_chunks(R, C)
    part = torch.empty(chunks * R, device=dev)
    rowsum = torch.empty(R, device=dev)
    diag = torch.zeros(R, device=dev) if want_diag else None
    with torch.cuda.device(dev):
        N.check(N.lib().cfm_infonce_rowsum(N.ptr(xb), N.ptr(yb), R, C, dp, temperature, diag_offset, N.ptr(rowsum),
                                           N.ptr(diag), N.ptr(part), N.stream_ptr()))
    return rowsum, diag


def infonce_rowcolsum(xb: torch.Tensor, yb: torch.Tensor, temperature: float, diag_offset: int = 0):
    """One pass over S = X.Y^T: rowsum[i] = sum_j exp((s_ij - 1)/T), colsum[j] = sum_i exp((s_ij - 1)/T) (the row sums
    of S^T) and diag[i] = s_(i, i+diag_offset)."""
    R, C, dp = xb.shape[0], yb.shape[0], xb.shape[1]
    dev = xb.device
    chunks = N.lib().cfm_simtile_chunks(R, C)
    part = torch.empty(chunks * R, device=dev)
    col_part = torch.empty(N.lib().cfm_infonce_colpart_floats(R, C), device=dev)
    rowsum, colsum = torch.empty(R, device=dev), torch.empty(C, device=dev)
    diag = torch.zeros(R, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().cfm_infonce_rowcolsum(N.ptr(xb), N.ptr(yb), R, C, dp, temperature, diag_offset, N.ptr(rowsum),
                                              N.ptr(colsum), N.ptr(diag), N.ptr(part), N.ptr(col_part), N.stream_ptr()))
    return rowsum, colsum, diag


def infonce_grad(xb, yb, D, temperature, diag_offset, B_total, rowsum_x, rowsum_y, diag, g_loss):
    R, C, dp = xb.shape[0], yb.shape[0], xb.shape[1]
    dev = xb.device
    chunks = N.lib().cfm_simtile_chunks(R, C)
    part = torch.empty(chunks * R * dp, device=dev)
    dx = torch.empty(R, D, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().cfm_infonce_grad(N.ptr(xb), N.ptr(yb), R, C, D, dp, temperature, diag_offset, B_total,
                                         N.ptr(rowsum_x), N.ptr(rowsum_y), N.ptr(diag), N.ptr(g_loss), N.ptr(dx), N.ptr(part),
                                         N.stream_ptr()))
    return dx


class InfoNCEFunction(torch.autograd.Function):
    """apply(firm_proj [B,D], ceo_proj [B,D], temperature) -> scalar symmetric InfoNCE loss.
    Three passes of the similarity-tile kernel (row AND column sums of S in one, then dF and dC); the [B,B] matrix
    only ever exists as 128x128 TMEM accumulators."""

    @staticmethod
    def forward(ctx, firm_proj, ceo_proj, temperature):
        _require_cuda(firm_proj, ceo_proj)
        B, D = firm_proj.shape
        fb, cb = pack_bf16(firm_proj), pack_bf16(ceo_proj)
        # one pass over S: its row sums (firm -> ceo) and its column sums (= row sums of S^T, ceo -> firm)
        rs_f, rs_c, diag = infonce_rowcolsum(fb, cb, temperature)
        loss = torch.empty((), device=firm_proj.device)
        with torch.cuda.device(firm_proj.device):
            N.check(N.lib().cfm_infonce_loss(N.ptr(rs_f), N.ptr(rs_c), N.ptr(diag), B, temperature, B, N.ptr(loss),
                                             N.stream_ptr()))
        ctx.save_for_backward(fb, cb, rs_f, rs_c, diag)
        ctx.temperature, ctx.D, ctx.in_dtype = temperature, D, firm_proj.dtype
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        fb, cb, rs_f, rs_c, diag = ctx.saved_tensors
        B = fb.shape[0]
        g = g_loss.contiguous().float()
        d_firm = infonce_grad(fb, cb, ctx.D, ctx.temperature, 0, B, rs_f, rs_c, diag, g)
        d_ceo = infonce_grad(cb, fb, ctx.D, ctx.temperature, 0, B, rs_c, rs_f, diag, g)
        return d_firm.to(ctx.in_dtype), d_ceo.to(ctx.in_dtype), None
