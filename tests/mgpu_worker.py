"""Worker for tests/test_gpu_multi.py: launched with torchrun (one process per GPU, NCCL).  Checks the data-parallel
two-tower step, InfoNCE with global negatives and the sharded top-k against the CPU oracle on the gathered batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ceo-recommender_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import oracle
    from helpers import assert_close_scaled, check_grads, load_into
    from ceo_firm_matching import CEOFirmMatcher, Config
    from ceo_firm_matching import distributed as D

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)

    # ---- data-parallel two-tower step: every replica ends with the gradient of the mean over the global batch ----
    f_cards, c_cards = [5000, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2]
    B = 300 if world <= 2 else 96
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=3)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    model = load_into(CEOFirmMatcher(meta, Config()), p).to(dev).train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.use_persistent_table_grads(True)
    dp = D.DataParallelTwoTower(model)
    def make_shards(seed):
        out = []
        for r in range(world):
            gen = torch.Generator().manual_seed(seed + r)
            out.append([torch.randn(B, 12, generator=gen),
                        torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1),
                        torch.randn(B, 2, generator=gen),
                        torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1),
                        torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5])
        return out

    specs = (oracle.two_tower_spec("firm"), oracle.two_tower_spec("ceo"))

    def clear_of_relu_kinks(seed0):
        # ReLU has no derivative at 0: of 60 candidate batches take the one whose hidden pre-activations stay furthest
        # from it (oracle.min_abs_relu_input), so that no unit's sign - a 1/B share of a column's gradients - hinges
        # on the last bits of the arithmetic (3xTF32 products resolve ~2e-6)
        best, best_margin = None, -1.0
        for seed in range(seed0, seed0 + 600, 10):
            cand = make_shards(seed)
            margin = min(oracle.min_abs_relu_input(p, specs, [(sh[0], sh[1]), (sh[2], sh[3])]) for sh in cand)
            if margin > best_margin:
                best, best_margin = cand, margin
        assert best_margin > 1.5e-5, best_margin
        return best

    shards = clear_of_relu_kinks(100)
    model.zero_grad_fast()
    loss, _ = model.forward_loss(*[t.to(dev) for t in shards[rank]])
    (loss * dp.loss_scale).backward()
    dp.sync_gradients()
    # oracle: average of the per-shard gradients (each shard with its own BatchNorm statistics)
    expected = None
    for r in range(world):
        po = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
        lo = oracle.weighted_mse(oracle.two_tower_forward(po, *shards[r][:4], training=True), shards[r][4], shards[r][5])
        lo.backward()
        g = {k: v.grad / world for k, v in po.items() if v.grad is not None}
        expected = g if expected is None else {k: expected[k] + g[k] for k in g}
    check_grads(model, expected, 1e-4, f"rank {rank} dp")
    flat = torch.cat([q.grad.reshape(-1) for q in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], x) for x in gathered), "replicas diverged"

    # ---- table-sharded two-tower step (tables owned by ranks, rows / gradient rows read over NVLink peer memory) ----
    model2 = load_into(CEOFirmMatcher(meta, Config()), p).to(dev).train()
    for m in model2.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model2.use_persistent_table_grads(True)
    ts = D.TableShardedTwoTower(model2, batch_rows=B)
    other = clear_of_relu_kinks(5000)        # different batches first: they touch rows (5000-row table) that the
    for step, data in enumerate((other, other, shards)):     # last step must find re-zeroed
        local = [t.to(dev) for t in data[rank]]
        model2.zero_grad_fast()
        if step > 0:                         # early publish: owners sort on a side stream while the step computes
            ts.begin_step(local[1], local[3])
        loss2, _ = model2.forward_loss(*local)
        (loss2 * ts.loss_scale).backward()
        ts.sync_gradients()
    assert torch.equal(loss2, loss), "stash forward differs from the direct gather"
    names = dict(model.named_parameters())
    owned_ids = {id(q) for q in ts.owned_parameters()}
    n_checked = 0
    for (t, k, c), r in zip(ts.slices, ts.owner):
        if r != rank:
            continue
        key = ("firm_embeddings" if t == 0 else "ceo_embeddings") + f".{k}.weight"
        w = names[key].shape[1] // ts.pieces[t]
        got = dict(model2.named_parameters())[key].grad[:, c * w:(c + 1) * w]
        assert id(dict(model2.named_parameters())[key]) in owned_ids
        assert torch.equal(got, names[key].grad[:, c * w:(c + 1) * w]), f"rank {rank}: slice {(t, k, c)} differs"
        n_checked += 1
    assert n_checked == sum(1 for r in ts.owner if r == rank)
    for k2, q in model2.named_parameters():
        if "embeddings" not in k2:
            assert_close_scaled(q.grad, names[k2].grad, 1e-6, f"sharded dense {k2}", floor=1e-7)
    ts.consolidate()

    # ---- InfoNCE with global negatives ----
    gen = torch.Generator().manual_seed(7)
    Bt, Dm = 1024 * world, 128
    f_all = F.normalize(torch.randn(Bt, Dm, generator=gen), dim=1)
    c_all = F.normalize(0.5 * f_all + torch.randn(Bt, Dm, generator=gen), dim=1)
    lo_, hi_ = D.shard_bounds(Bt, world, rank)
    f = f_all[lo_:hi_].to(dev).requires_grad_(True)
    c = c_all[lo_:hi_].to(dev).requires_grad_(True)
    loss = D.info_nce_loss_global(f, c, 0.07)
    loss.backward()
    fq = f_all.bfloat16().float().requires_grad_(True)
    cq = c_all.bfloat16().float().requires_grad_(True)
    lref = oracle.info_nce(fq, cq, 0.07)
    lref.backward()
    assert abs(float(loss) - float(lref)) <= 2e-5 * abs(float(lref)) + 1e-5, (float(loss), float(lref))
    assert_close_scaled(f.grad, fq.grad[lo_:hi_], 6e-3, "global d_firm")
    assert_close_scaled(c.grad, cq.grad[lo_:hi_], 6e-3, "global d_ceo")

    # ---- row-sharded all-pairs top-k with cross-shard merge ----
    rows_all = F.normalize(torch.randn(2000, 60, generator=gen), dim=1)
    cols_all = F.normalize(torch.randn(30001, 60, generator=gen), dim=1)       # ragged column shards
    rlo, rhi = D.shard_bounds(2000, world, rank)
    clo, chi = D.shard_bounds(30001, world, rank)
    so, io = oracle.allpairs_topk(rows_all[rlo:rhi], cols_all, 100, 14.2857)
    for how in ("fused", "shards"):
        s, i = D.score_topk_sharded(rows_all[rlo:rhi].to(dev), cols_all[clo:chi].to(dev), 100, 14.2857, merge=how)
        np.testing.assert_array_equal(i.cpu().numpy(), io.numpy())
        assert_close_scaled(s, so, 2e-6, f"sharded scores ({how})")

    dist.barrier()
    if rank == 0:
        print("MULTIGPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
