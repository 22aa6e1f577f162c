"""Graph-replayed sharded step under ablations (timing only; ablated variants are NOT correct)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching import distributed as D, ops
from ceo_firm_matching.training import GraphedTwoTowerStep
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
batches = bench.make_batches(8, bench.B_PER_GPU, dev, 1234 + rank)

def run(tag, local_tables=False, no_reduce=False, no_begin=False, no_sync=False):
    model = bench.build_model(dev, "fp32")
    ts = D.TableShardedTwoTower(model, batch_rows=bench.B_PER_GPU)
    if local_tables:
        for h in model._handles:
            h.row_source.tables = [e.weight for e in h.embeddings for _ in range(h.row_source.pieces)]
    if no_reduce:
        ts.kernels.peer_reduce = lambda *a, **k: None
        ts.kernels.rezero = lambda *a, **k: None
    runner = GraphedTwoTowerStep(model, batches[0], optimizer=None, warmup=3, loss_scale=ts.loss_scale,
                                 after_backward=None if no_sync else ts.sync_gradients,
                                 before_forward=None if no_begin else ts.begin_step)
    for i in range(5): runner.step(batches[i % 8])
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): runner.step(batches[i % 8])
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 40], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{tag:34s} {float(ms):.4f} ms/step", flush=True)
    runner.graph.reset(); torch.cuda.synchronize()

run("full")
run("local tables (no NVLink gather)", local_tables=True)
run("no owner reduce", no_reduce=True)
run("no begin_step (sort after bwd)", no_begin=True)
run("no sync at all", no_begin=True, no_sync=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
