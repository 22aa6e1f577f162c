"""Phase times of the table-sharded step (eager, CUDA events on rank 0)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching import distributed as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
model = bench.build_model(dev, "fp32")
ts = D.TableShardedTwoTower(model, batch_rows=bench.B_PER_GPU)
batches = bench.make_batches(8, bench.B_PER_GPU, dev, 1234 + rank)
names = ["zero", "begin", "fwd", "bwd", "sync"]
acc = {k: 0.0 for k in names}
side = torch.cuda.Stream()
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
with torch.cuda.stream(side):
    for it in range(25):
        b = batches[it % 8]
        e = [ev()]
        model.zero_grad_fast(); e.append(ev())
        ts.begin_step(b[1], b[3]); e.append(ev())
        loss, _ = model.forward_loss(*b); e.append(ev())
        (loss * ts.loss_scale).backward(); e.append(ev())
        ts.sync_gradients(); e.append(ev())
        torch.cuda.synchronize()
        if it >= 5:
            for i, k in enumerate(names):
                acc[k] += e[i].elapsed_time(e[i + 1]) / 20
    # NCCL latency of a tiny all-reduce, back to back
    tok = torch.zeros(1, device=dev)
    for _ in range(5): dist.all_reduce(tok)
    torch.cuda.synchronize(); a = ev()
    for _ in range(50): dist.all_reduce(tok)
    b_ = ev(); torch.cuda.synchronize()
if rank == 0:
    print("phases ms", {k: round(v, 4) for k, v in acc.items()}, "sum", round(sum(acc.values()), 4))
    print("tiny all-reduce us", a.elapsed_time(b_) / 50 * 1e3)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
