"""world_size-2 gloo tests (CPU) of the multi-GPU protocol in ceo_firm_matching.distributed: sharding, bucketed
all-reduce, ragged gathers, InfoNCE with global negatives and the sharded top-k merge.  The per-rank math is
injected as a torch reference backend (the CUDA kernels need a GPU); the collectives and bookkeeping are the
product's own."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class TorchInfoNCEBackend:
    """Reference math with the same contract as CudaInfoNCEBackend (fixed maximum 1/T, fp32)."""

    def pack(self, x):
        return x.detach().float()

    def rowsum(self, xb, yb, temperature, diag_offset, want_diag):
        s = xb @ yb.t()
        rs = torch.exp((s - 1) / temperature).sum(1)
        diag = s[torch.arange(xb.shape[0]), torch.arange(xb.shape[0]) + diag_offset] if want_diag else None
        return rs, diag

    def local_loss(self, rs_row, rs_col, diag, temperature, b_total):
        return (torch.log(rs_row) + torch.log(rs_col) + 2 / temperature - 2 * diag / temperature).sum() / (2 * b_total)

    def grad(self, xb, yb, d, temperature, diag_offset, b_total, rs_x, rs_y, diag, g_loss):
        e = torch.exp((xb @ yb.t() - 1) / temperature)
        g = e * (1 / rs_x[:, None] + 1 / rs_y[None, :]) / (2 * b_total * temperature)
        idx = torch.arange(xb.shape[0])
        g[idx, idx + diag_offset] -= 1 / (b_total * temperature)
        return g_loss * (g @ yb)[:, :d]


class TorchScoringBackend:
    def topk(self, rows, cols, k, scale, col_offset):
        s = rows.double() @ cols.double().t()
        order = torch.argsort(-s, dim=1, stable=True)[:, :k]
        out_s = torch.full((rows.shape[0], k), float("-inf"))
        out_i = torch.full((rows.shape[0], k), -1, dtype=torch.int64)
        out_s[:, :order.shape[1]] = (torch.gather(s, 1, order) * scale).float()
        out_i[:, :order.shape[1]] = order + col_offset
        return out_s, out_i

    def merge(self, ps, pi):
        n_parts, R, k = ps.shape
        s = ps.permute(1, 0, 2).reshape(R, -1).double()
        i = pi.permute(1, 0, 2).reshape(R, -1)
        s = torch.where(i < 0, torch.full_like(s, float("-inf")), s)
        key = torch.argsort(i, dim=1, stable=True)                           # index asc ...
        s, i = torch.gather(s, 1, key), torch.gather(i, 1, key)
        order = torch.argsort(-s, dim=1, stable=True)[:, :k]                  # ... then score desc, stable
        return torch.gather(s, 1, order).float(), torch.gather(i, 1, order)


class GlooPeers:
    """Stand-in for CudaIpcPeers on CPU: there is no shared memory, so every rank keeps a copy of each published
    buffer and ``fence()`` refreshes the copies from their publishers."""

    def share(self, named):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        metas = [None] * self.world
        dist.all_gather_object(metas, {k: (tuple(v.shape), v.dtype) for k, v in named.items()})
        self.views = {}
        for r in range(self.world):
            for k, (shape, dtype) in metas[r].items():
                self.views.setdefault(k, [None] * self.world)[r] = (named[k] if r == self.rank
                                                                    else torch.zeros(shape, dtype=dtype))
        self.fence()
        return self.views

    def fence(self):
        for k in sorted(self.views):
            for r, t in enumerate(self.views[k]):
                if t is not None:
                    dist.broadcast(t, src=r)


class TorchTableOps:
    """Reference math with the contract of CudaTableOps (the row source is only used by the CUDA forward)."""

    def make_row_source(self, handle, tables, pieces, dx_emb):
        return ("row-source", len(tables), pieces)

    def make_scratch(self, n_owned, n_peers, B, device):
        return {"touched": []}

    def peer_reduce(self, groups, n_peers, B, scratch, phase=0):
        if phase != 2:                                                 # "keys + sort": needs only the indices
            scratch["idx"] = [[[o["x_cat"][r][:, o["col"]].clone() for r in range(n_peers)] for o in g["owned"]]
                              for g in groups]
        if phase == 1:
            return
        scratch["touched"] = []
        for g, gi in zip(groups, scratch["idx"]):
            E, width = g["emb_dim"], g["width"]
            for o, oi in zip(g["owned"], gi):
                c0 = o["col0"]
                for r in range(n_peers):                               # rank order, then row order
                    rows = o["dx_emb"][r][:, o["col"] * E + c0:o["col"] * E + c0 + width]
                    o["grad"][:, c0:c0 + width].index_add_(0, oi[r], rows)
                    scratch["touched"].append((o, oi[r], width))

    def rezero(self, groups, n_peers, B, scratch):
        for o, idx, width in scratch["touched"]:
            o["grad"][idx, o["col0"]:o["col0"] + width] = 0.0

    def side_stream(self, device):
        return None


def _table_sharded_protocol(rank, world):
    """TableShardedTwoTower bookkeeping end to end with emulated peers: plan, consolidation, dense all-reduce,
    owner-side reduce over every rank's (index, gradient row) pairs, sparse re-zero on the next step."""
    from ceo_firm_matching import CEOFirmMatcher, Config
    from ceo_firm_matching import distributed as D
    f_cards, c_cards, B = [40, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2], 16
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    torch.manual_seed(100 + rank)                                      # replicas start DIFFERENT on purpose
    model = CEOFirmMatcher(meta, Config())
    model.use_persistent_table_grads(True)
    ts = D.TableShardedTwoTower(model, batch_rows=B, peers=GlooPeers(), table_ops=TorchTableOps())
    hs = model._handles
    res = {"plan": (ts.pieces, ts.slices, ts.owner)}
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    res["consolidated"] = all(torch.equal(both[0], x) for x in both)    # dense from rank 0, slices from owners
    expected_steps = []
    for step in range(3):
        for p in ts.dense:
            p.grad = torch.full_like(p, float(rank + 1 + step))
        gens = [torch.Generator().manual_seed(1000 * step + r) for r in range(world)]
        data = [[(torch.stack([torch.randint(0, e.num_embeddings, (B,), generator=g) for e in h.embeddings], 1),
                  torch.randn(B, h.n_tables * h.emb_dim, generator=g)) for h in hs] for g in gens]
        if step == 2:                                                   # early publish + sort, reduce after backward
            ts.begin_step(data[rank][0][0], data[rank][1][0])
        for t, h in enumerate(hs):
            h.table_grads.pending = data[rank][t]                       # what the CUDA backward leaves behind
        ts.sync_gradients()
        ok = all(torch.equal(p.grad, torch.full_like(p, float(sum(r + 1 + step for r in range(world)))))
                 for p in ts.dense)
        for (t, k, c), r in zip(ts.slices, ts.owner):
            h = hs[t]
            w = h.emb_dim // ts.pieces[t]
            exp = torch.zeros(h.embeddings[k].num_embeddings, w)
            for rr in range(world):
                x, dx = data[rr][t]
                exp.index_add_(0, x[:, k], dx[:, k * h.emb_dim + c * w:k * h.emb_dim + (c + 1) * w])
            got = h.embeddings[k].weight.grad[:, c * w:(c + 1) * w]
            ok = ok and (torch.equal(got, exp) if r == rank else float(got.abs().sum()) == 0.0)
        expected_steps.append(ok)
    res["steps"] = expected_steps
    owned = {id(p) for p in ts.owned_parameters()}
    res["owned_ok"] = all((id(hs[t].embeddings[k].weight) in owned) == (r == rank) or ts.pieces[t] > 1
                          for (t, k, c), r in zip(ts.slices, ts.owner))
    return res


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "ceo-recommender_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from ceo_firm_matching import distributed as D
        res = {}
        # --- shard bookkeeping + bucketed all-reduce ---
        assert [D.shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
        a, b = torch.full((3, 2), float(rank + 1)), torch.arange(4.0) * (rank + 1)
        D.allreduce_flat_([a, None, b], scale=0.5)
        res["allreduce"] = (a.clone(), b.clone())
        rag, counts = D.gather_ragged_rows(torch.full((rank + 1, 2), float(rank)))
        res["ragged"] = (rag.clone(), counts)
        # --- InfoNCE with global negatives ---
        gen = torch.Generator().manual_seed(0)
        Bt, Dm = 12, 16
        f_all = F.normalize(torch.randn(Bt, Dm, generator=gen), dim=1)
        c_all = F.normalize(0.5 * f_all + torch.randn(Bt, Dm, generator=gen), dim=1)
        lo, hi = D.shard_bounds(Bt, world, rank)
        f = f_all[lo:hi].clone().requires_grad_(True)
        c = c_all[lo:hi].clone().requires_grad_(True)
        loss = D.info_nce_loss_global(f, c, 0.07, backend=TorchInfoNCEBackend())
        (2.0 * loss).backward()
        fo, co = f_all.clone().requires_grad_(True), c_all.clone().requires_grad_(True)
        lref = oracle.info_nce(fo, co, 0.07)
        (2.0 * lref).backward()
        res["nce"] = (float(loss), float(lref), (f.grad - fo.grad[lo:hi]).abs().max().item(),
                      (c.grad - co.grad[lo:hi]).abs().max().item())
        # --- sharded scoring with ragged shards + merge ---
        rows_all = F.normalize(torch.randn(9, 8, generator=gen), dim=1)
        cols_all = F.normalize(torch.randn(31, 8, generator=gen), dim=1)
        cols_all[20] = cols_all[3]                                             # an exact tie across shards
        rlo, rhi = D.shard_bounds(9, world, rank)
        clo, chi = (0, 20) if rank == 0 else (20, 31)
        so, io = oracle.allpairs_topk(rows_all[rlo:rhi], cols_all, 7, 3.0)
        same, err = True, 0.0
        for how in ("shards", "fused"):
            s, i = D.score_topk_sharded(rows_all[rlo:rhi], cols_all[clo:chi], 7, 3.0, backend=TorchScoringBackend(),
                                        merge=how)
            same, err = same and torch.equal(i, io), max(err, (s - so).abs().max().item())
        res["topk"] = (same, err)
        res["sharded"] = _table_sharded_protocol(rank, world)
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_world2_protocol():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        r = out[rank]
        a, b = r["allreduce"]
        assert torch.equal(a, torch.full((3, 2), 1.5)) and torch.equal(b, torch.arange(4.0) * 1.5)
        rag, counts = r["ragged"]
        assert counts == [1, 2] and torch.equal(rag, torch.tensor([[0., 0.], [1., 1.], [1., 1.]]))
        loss, lref, ef, ec = r["nce"]
        assert loss == pytest.approx(lref, rel=1e-5) and ef < 1e-5 and ec < 1e-5
        same_idx, err = r["topk"]
        assert same_idx and err < 1e-5
    assert out[0]["nce"][0] == out[1]["nce"][0]                     # the global loss is identical on every rank
    for rank in range(world):
        sh = out[rank]["sharded"]
        assert sh["plan"] == out[0]["sharded"]["plan"]              # every rank derives the same ownership
        assert sh["consolidated"] and sh["steps"] == [True, True, True] and sh["owned_ok"]


def test_table_slice_plan_is_balanced():
    sys.path.insert(0, os.path.join(ROOT, "ceo-recommender_b200"))
    from ceo_firm_matching.distributed import plan_table_slices
    towers = [(4, 48), (7, 8)]                                      # BASELINE config 4: 992 B of rows per pair
    for world in (1, 2, 4, 8):
        pieces, slices, owner = plan_table_slices(towers, world)
        load = [0] * world
        for (t, k, c), r in zip(slices, owner):
            load[r] += towers[t][1] // pieces[t]
        assert sum(load) == 4 * 48 + 7 * 8 and len(slices) == 4 * pieces[0] + 7 * pieces[1]
        assert max(load) <= 1.06 * (sum(load) / world) + 8, (world, pieces, load)
        assert all((towers[t][1] // pieces[t]) % 4 == 0 for t in range(2))      # rows stay 16-byte vectors
    assert plan_table_slices(towers, 8)[0] == [2, 1]                # wide firm tables are cut in two at 8 ranks
