"""Where does the end-to-end loop lose time?  Variants of bench.py's e2e loop."""
import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
from ceo_firm_matching.training import GraphedTwoTowerStep
dev = torch.device('cuda', 0)
model = bench.build_model(dev, "fp32")
batches = bench.make_batches(8, bench.B_PER_GPU, dev, 1234)
runner = GraphedTwoTowerStep(model, batches[0], optimizer=None, warmup=3)
host = [[t.cpu().pin_memory() for t in b] for b in batches[:4]]
copy_stream = torch.cuda.Stream(dev)
loss_host = torch.zeros(1).pin_memory()
K = 40

def timeit(fn, label):
    fn(3); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); fn(K); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"{label:60s} {best / K * 1e3:.3f} ms/step", flush=True)

def dev_only(n):
    for i in range(n): runner.step(batches[i % 8])
timeit(dev_only, "device-resident batches (no H2D)")

def h2d_only(n):
    with torch.cuda.stream(copy_stream):
        for i in range(n):
            x = [t.to(dev, non_blocking=True) for t in host[i % 4]]
    copy_stream.synchronize()
timeit(h2d_only, "H2D alone (6 tensors, 10 MB)")

def upload(i):
    with torch.cuda.stream(copy_stream):
        d = [t.to(dev, non_blocking=True) for t in host[i % 4]]
        ev = torch.cuda.Event(); ev.record(copy_stream)
    return d, ev
def as_bench(n):
    nxt = upload(0)
    for i in range(n):
        d, ev = nxt
        if i + 1 < n: nxt = upload(i + 1)
        torch.cuda.current_stream().wait_event(ev)
        loss = runner.step(d)
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        for t in d: t.record_stream(torch.cuda.current_stream())
timeit(as_bench, "bench.py loop (default stream)")

side = torch.cuda.Stream(dev)
def on_side(n):
    with torch.cuda.stream(side):
        as_bench(n)
timeit(on_side, "same loop on a non-default stream")

# H2D straight into the graph's static inputs, double-buffer-free: copy stream writes static buffers after the
# previous replay has consumed them (event), replay waits for the copy
done = torch.cuda.Event(); done.record()
def direct(n):
    global done
    for i in range(n):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)                     # previous replay finished reading the static inputs
            for s, t in zip(runner.static, host[i % 4]): s.copy_(t, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        torch.cuda.current_stream().wait_event(ev)
        loss = runner.step(None)
        done = torch.cuda.Event(); done.record()
        loss_host.copy_(loss.reshape(1), non_blocking=True)
timeit(direct, "H2D directly into static inputs (serialised with the step)")
