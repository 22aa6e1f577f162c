import os, sys, torch, torch.distributed as dist
ROOT = '/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/ceo-recommender_b200'); sys.path.insert(0, ROOT + '/tests')
import oracle
from helpers import load_into
from ceo_firm_matching import CEOFirmMatcher, Config
from ceo_firm_matching import distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
f_cards, c_cards, B = [5000, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2], 300
p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=3)
meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
model = load_into(CEOFirmMatcher(meta, Config()), p).to(dev).train()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
model.use_persistent_table_grads(True)
dp = D.DataParallelTwoTower(model)
shards = []
for r in range(world):
    gen = torch.Generator().manual_seed(100 + r)
    shards.append([torch.randn(B, 12, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1),
                   torch.randn(B, 2, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1),
                   torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5])
model.zero_grad_fast()
loss, preds = model.forward_loss(*[t.to(dev) for t in shards[rank]])
(loss * dp.loss_scale).backward()
pend = model._handles[0].table_grads.pending
dx_local = pend[1].clone()
dp.sync_gradients()
po = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in p.items()}
pr = oracle.two_tower_forward(po, *shards[rank][:4], training=True)
lo = oracle.weighted_mse(pr, shards[rank][4], shards[rank][5]); lo.backward()
if rank == 0:
    print("loss", float(loss), float(lo), "preds maxdiff", float((preds.cpu() - pr).abs().max()))
    g = model.firm_embeddings[0].weight.grad.cpu()
    eo = po["firm_embeddings.0.weight"].grad / world
    x0 = shards[0][1][:, 0]; x1 = shards[1][1][:, 0]
    # local contribution check: rows only touched by rank 0
    only0 = [int(i) for i in set(x0.tolist()) - set(x1.tolist())]
    d = (g[only0] - eo[only0]).abs().max()
    print("rows only rank0 touches: maxdiff", float(d), "n", len(only0))
    # dx_emb local vs oracle per-pair rows: oracle row grad for unique idx
    uniq = [i for i in only0 if (x0 == i).sum() == 1]
    b = [int((x0 == i).nonzero()[0]) for i in uniq]
    print("dx_local rows vs oracle", float((dx_local.cpu()[b, :48] - eo[uniq]).abs().max()))
    bad = ((g - eo).abs() > 1e-5).nonzero()
    print("bad entries", bad.shape[0], bad[:10].tolist())
    print("row 1237: in x0", int((x0 == 1237).sum()), "in x1", int((x1 == 1237).sum()))
dist.barrier(); dist.destroy_process_group()
