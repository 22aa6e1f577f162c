"""Single-GPU parity of the building blocks of the NVLink table-sharded mode (distributed.TableShardedTwoTower):
the row stash must reproduce the direct gather bit for bit, and the owner-side peer reduce must equal the ordinary
segment reduce of the concatenated global batch bit for bit.  (The cross-GPU run is tests/test_gpu_multi.py.)"""
import pytest
import torch

import oracle
from helpers import load_into

pytestmark = pytest.mark.gpu
DEV = "cuda"
F_CARDS, C_CARDS = [50, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2]


def _model(seed=4):
    from ceo_firm_matching import CEOFirmMatcher, Config
    p = oracle.init_two_tower_params(12, F_CARDS, 2, C_CARDS, seed=seed)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": F_CARDS, "n_ceo_numeric": 2, "ceo_cat_counts": C_CARDS}
    return load_into(CEOFirmMatcher(meta, Config()), p).to(DEV).train()


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    return [x.to(DEV) for x in (
        torch.randn(B, 12, generator=g), torch.stack([torch.randint(0, n, (B,), generator=g) for n in F_CARDS], 1),
        torch.randn(B, 2, generator=g), torch.stack([torch.randint(0, n, (B,), generator=g) for n in C_CARDS], 1),
        torch.randn(B, 1, generator=g), torch.rand(B, 1, generator=g) + 0.5)]


def _sliced_tables(h, pieces):
    return [e.weight for e in h.embeddings for _ in range(pieces)]


@pytest.mark.parametrize("B,pieces", [(1, 1), (77, 2), (1000, 1), (1000, 4)])
def test_gather_rows_matches_index_select(B, pieces):
    from ceo_firm_matching import ops
    m = _model()
    for h, col in zip(m._handles, (1, 3)):
        if (h.emb_dim // pieces) * pieces != h.emb_dim:
            continue
        x_cat = _batch(max(B, 2), 5)[col][:B].contiguous()
        stash, index = ops.StashedRows(h, _sliced_tables(h, pieces), pieces).gather(x_cat, B)
        assert stash.shape == (h.n_tables, B, h.emb_dim) and index.shape == (B, h.n_tables)
        for k, e in enumerate(h.embeddings):
            assert torch.equal(stash[k], e.weight[x_cat[:, k]])
            assert torch.equal(index[:, k], torch.arange(B, device=DEV))


def test_gather_rows_reads_each_slice_from_its_own_tensor():
    """Column slice c of a table must come from tables[k*pieces + c] (its owner's copy), not from slice 0's."""
    from ceo_firm_matching import ops
    m = _model()
    h = m._handles[0]
    pieces, w = 2, h.emb_dim // 2
    copies = [[e.weight.detach().clone() + 100.0 * c for c in range(pieces)] for e in h.embeddings]
    x_cat = _batch(50, 5)[1]
    stash, _ = ops.StashedRows(h, [t for per in copies for t in per], pieces).gather(x_cat, 50)
    for k in range(h.n_tables):
        for c in range(pieces):
            assert torch.equal(stash[k][:, c * w:(c + 1) * w], copies[k][c][x_cat[:, k], c * w:(c + 1) * w])


def test_gather_rows_flags_out_of_range_index():
    from ceo_firm_matching import ops
    m = _model()
    h = m._handles[0]
    x_cat = _batch(8, 5)[1]
    x_cat[3, 1] = 999
    ops.StashedRows(h, _sliced_tables(h, 1)).gather(x_cat, 8)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        ops.raise_if_index_error(torch.device(DEV))


@pytest.mark.parametrize("pieces", [1, 2])
def test_stash_mode_is_bitwise_the_direct_path(pieces):
    from ceo_firm_matching import ops
    batch = _batch(300, 6)
    results = []
    for stash in (False, True):
        m = _model()
        for mod in m.modules():                             # the dropout counter advances per call: switch it off
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        if stash:
            for h in m._handles:
                h.row_source = ops.StashedRows(h, _sliced_tables(h, pieces), pieces)
        loss, preds = m.forward_loss(*batch)
        loss.backward()
        results.append([loss.detach(), preds.detach()] + [p.grad for p in m.parameters()])
    for a, b in zip(*results):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n_peers,pieces", [(1, 1), (2, 1), (3, 2)])
def test_peer_reduce_equals_segment_reduce_of_concatenated_batch(n_peers, pieces):
    """`n_peers` ranks' buffers emulated on one device; the owner holds slices of tables from both towers."""
    from ceo_firm_matching import distributed as D, ops
    B = 257
    m = _model()
    m.use_persistent_table_grads(True)
    hs = m._handles
    xs = [[_batch(B, 20 + r)[1], _batch(B, 20 + r)[3]] for r in range(n_peers)]
    g = torch.Generator().manual_seed(3)
    dxs = [[torch.randn(B, h.n_tables * h.emb_dim, generator=g).to(DEV) for h in hs] for r in range(n_peers)]
    # expected: the replicated data-parallel reduce over the concatenated batch (rank order)
    for hi, h in enumerate(hs):
        ops.reduce_table_grads(h, torch.cat([xs[r][hi] for r in range(n_peers)]),
                               torch.cat([dxs[r][hi] for r in range(n_peers)]))
    kernels = D.CudaTableOps()
    groups = []
    for hi, cols in ((0, [1, 3]), (1, [0, 2, 6])):
        h = hs[hi]
        w = h.emb_dim // pieces
        if w * pieces != h.emb_dim or w % 4:
            w = h.emb_dim                                      # this tower keeps whole tables
        owned = [dict(n_cols=h.n_tables, col=k, col0=c * w, grad=torch.zeros_like(h.embeddings[k].weight),
                      x_cat=[xs[r][hi] for r in range(n_peers)], dx_emb=[dxs[r][hi] for r in range(n_peers)], tower=hi)
                 for k in cols for c in range(h.emb_dim // w) if (k + c) % 2 == 0 or w == h.emb_dim]
        groups.append(dict(owned=owned, emb_dim=h.emb_dim, width=w))
    n_owned = sum(len(g["owned"]) for g in groups)
    scratch = kernels.make_scratch(n_owned, n_peers, B, torch.device(DEV))
    for phases in ((0,), (1, 2)):                               # one call, or sort first and reduce later
        for ph in phases:
            kernels.peer_reduce(groups, n_peers, B, scratch, phase=ph)
        for g in groups:
            for o in g["owned"]:
                w = g["width"]
                exp = torch.zeros_like(o["grad"])
                exp[:, o["col0"]:o["col0"] + w] = hs[o["tower"]].embeddings[o["col"]].weight.grad[:, o["col0"]:o["col0"] + w]
                assert torch.equal(o["grad"], exp), (o["tower"], o["col"], o["col0"])
        kernels.rezero(groups, n_peers, B, scratch)
        assert all(float(o["grad"].abs().sum()) == 0.0 for g in groups for o in g["owned"])


def test_joint_reduce_is_bitwise_the_per_tower_reduce():
    """model.use_persistent_table_grads(True) sends both towers' (table, index) pairs through ONE radix sort
    (cfm_emb_grad_joint_reduce); the dense table gradients must equal the per-tower autograd path bit for bit,
    also on a second step with other indices (sparse re-zero of the rows the first step touched)."""
    ma, mb = _model(), _model()
    for m in (ma, mb):
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    ma.use_persistent_table_grads(True)
    assert ma._handles[0].table_grads.joint is ma._handles[1].table_grads.joint is not None
    for seed, B in ((6, 300), (7, 300), (8, 77)):
        batch = _batch(B, seed)
        ma.zero_grad_fast()
        mb.zero_grad(set_to_none=True)
        la, _ = ma.forward_loss(*batch)
        la.backward()
        lb, _ = mb.forward_loss(*batch)
        lb.backward()
        for (ka, pa), (kb, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            assert ka == kb and torch.equal(pa.grad, pb.grad), ka
