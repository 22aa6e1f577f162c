#!/usr/bin/env python
"""Headline benchmark: two-tower training step (fwd + weighted-MSE + bwd) on BASELINE config 4.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one synthetic batch of 65,536 CEO-firm pairs per GPU with
1M-row categorical embedding tables (4 firm x 48 + 7 CEO x 8 floats), dropout on, dense embedding
gradients (persistent buffers, rows re-zeroed sparsely).  The optimiser step is NOT part of the metric
 and is reported separately.  Prints ONE JSON line (see the task contract).

``--impl reference`` times the CPU restatement of the reference (oracle/, torch CPU ops == the reference's
own ATen path) on the box's host cores for the same config; ``/root/reference`` itself cannot travel.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ceo-recommender_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

B_PER_GPU = 65_536
N_PAIRS = 10_000_000
TABLE_ROWS = 1_000_000
F_CARDS, C_CARDS = [TABLE_ROWS] * 4, [TABLE_ROWS] * 7
BYTES_PER_PAIR = 2228          # 1148 fwd + 1080 bwd algorithmic bytes, large-table regime
BYTES_BWD1_PER_PAIR = 1080     # stage-1 backward kernel: re-read indices (88) + emit embedding-grad rows (992)
METRIC = "two_tower_train_pairs_per_sec"
WORKLOADS = {
    "fp32": "config4: two-tower fwd+loss+bwd, B=65536/GPU, 10M synthetic pairs, 11 tables x 1M rows, fp32 end to end",
    "tf32": ("config4: two-tower fwd+loss+bwd, B=65536/GPU, 10M synthetic pairs, 11 tables x 1M rows, "
             "single-pass TF32 tensor-core tower products with fp32 accumulation, everything else fp32"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return d["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------
# synthetic data of config 4's shape
# ----------------------------------------------------------------------------------------------
def make_batches(n_batches, B, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for _ in range(n_batches):
        f_num = torch.randn(B, 12, device=device, generator=g)
        c_num = torch.randn(B, 2, device=device, generator=g)
        f_cat = torch.randint(0, TABLE_ROWS, (B, 4), device=device, generator=g)
        c_cat = torch.randint(0, TABLE_ROWS, (B, 7), device=device, generator=g)
        target = torch.randn(B, 1, device=device, generator=g)
        sd = torch.rand(B, 1, device=device, generator=g) * 0.9 + 0.1
        out.append((f_num, f_cat, c_num, c_cat, target, 1.0 / (sd * sd + 1e-6)))
    return out


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML polling thread (2 ms period;
    the timed region of the default run is ~25 ms, shorter than one `nvidia-smi -lms` tick), nvidia-smi as fallback.
    `mark()` / `unmark()` bracket the timed region; samples outside it are kept separately."""

    def __init__(self, index):
        self.index, self.samples, self.stop, self.timed, self.thread, self.max_mhz = index, [], False, False, None, None
        self.backend = None

    def _nvml_loop(self, nv, handle):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                self.samples.append((self.timed, float(mhz), [k for k, b in bits.items() if mask & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                     "--format=csv,noheader,nounits", "-lms", "50"],
                                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.backend = "unavailable"                   # neither NVML nor nvidia-smi: no samples, say so
            return
        self.proc = proc
        for ln in proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            try:
                self.max_mhz = float(f[1])
                self.samples.append((self.timed, float(f[0]),
                                     [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")]))
            except Exception:
                continue
            if self.stop:
                break

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.backend = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
        except Exception:
            self.backend = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
        self.thread.start()
        return self

    def mark(self):
        self.timed = True

    def unmark(self):
        self.timed = False

    def __exit__(self, *a):
        self.stop = True
        if getattr(self, "proc", None) is not None:
            self.proc.terminate()
        self.thread.join(timeout=2)

    def summary(self):
        timed = [s for s in self.samples if s[0]]
        use = timed if timed else self.samples            # fallback sampler may tick slower than the timed region
        sm = [s[1] for s in use]
        reasons = sorted({r for s in use for r in s[2]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": self.backend,
                "window": "timed region" if timed else "step loop around the timed region"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (torch CPU restatement of the reference) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_step_fn():
    import oracle
    torch.set_num_threads(os.cpu_count())
    p = oracle.init_two_tower_params(12, F_CARDS, 2, C_CARDS, seed=0)
    names = [k for k, v in p.items() if v.is_floating_point() and "running" not in k]
    for k in names:
        p[k].requires_grad_(True)
    gen = torch.Generator().manual_seed(1234)
    masks_p = 0.1

    def batch():
        B = B_PER_GPU
        return (torch.randn(B, 12, generator=gen), torch.randint(0, TABLE_ROWS, (B, 4), generator=gen),
                torch.randn(B, 2, generator=gen), torch.randint(0, TABLE_ROWS, (B, 7), generator=gen),
                torch.randn(B, 1, generator=gen), 1.0 / (torch.rand(B, 1, generator=gen) * 0.9 + 0.1) ** 2)

    data = [batch() for _ in range(2)]

    def step(i):
        f_num, f_cat, c_num, c_cat, target, weights = data[i % len(data)]
        for k in names:
            p[k].grad = None
        B = f_num.shape[0]
        masks = {s: [torch.rand(B, w, generator=gen) >= masks_p for w in (64, 32)] for s in ("firm", "ceo")}
        preds = oracle.two_tower_forward(p, f_num, f_cat, c_num, c_cat, training=True, masks=masks)
        loss = oracle.weighted_mse(preds, target, weights)
        loss.backward()
        return float(loss)

    return step


def run_cpu(steps, warmup):
    step = cpu_step_fn()
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = time.perf_counter() - t0
    return B_PER_GPU * steps / dt, dt / steps * 1e3


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms = run_cpu(args.steps, max(args.warmup, 1))
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.precision], "timing": "host perf_counter, inputs in host memory"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of one B=65536 batch (fwd+loss+bwd, dense 1M-row table grads), "
                                   "oracle = torch-CPU restatement of the reference"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def build_model(device, precision="fp32"):
    from ceo_firm_matching import CEOFirmMatcher, Config
    torch.manual_seed(0)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": F_CARDS, "n_ceo_numeric": 2, "ceo_cat_counts": C_CARDS}
    model = CEOFirmMatcher(meta, Config()).to(device).train()
    model.set_precision(precision)
    model.use_persistent_table_grads(True)
    return model


def _time_cuda(fn, reps, dist=None, device=None):
    """Best-of-`reps` device time of fn() in ms (CUDA events on the current stream; max over ranks)."""
    best = float("inf")
    for _ in range(reps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = min(best, ms)
    return best


def secondary_metrics(device, world, rank, dist):
    """The other two paths the metric names: in-batch InfoNCE (config 3: global batch 65,536, D=128, bf16) and
    all-pairs scoring with top-100 (config 5: 1M x 1M, D=60), sharded by rows when world > 1."""
    import torch.nn.functional as F
    from ceo_firm_matching.contrastive import info_nce_loss
    from ceo_firm_matching.scoring import score_topk
    from ceo_firm_matching import distributed as D
    try:
        bf16_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        bf16_peak = 1590.0
    out = {}
    g = torch.Generator(device=device).manual_seed(100 + rank)
    # ---- config 3 ----
    Bg, Dm = 65_536, 128
    n = Bg // world
    f = F.normalize(torch.randn(n, Dm, device=device, generator=g), dim=1).requires_grad_(True)
    c = F.normalize(torch.randn(n, Dm, device=device, generator=g), dim=1).requires_grad_(True)

    def nce_step():
        f.grad = c.grad = None
        loss = info_nce_loss(f, c, 0.07) if world == 1 else D.info_nce_loss_global(f, c, 0.07)
        loss.backward()

    nce_step()
    ms = _time_cuda(nce_step, 3, dist, device)
    flops = 6.0 * Bg * Bg * Dm                       # fwd 2 B^2 D + bwd 4 B^2 D (recompute not credited)
    out["infonce_fwd_bwd"] = {"workload": "config3: B=65536 global, D=128, bf16 operands, fp32 accumulate",
                              "ms": ms, "value": Bg * Bg / (ms * 1e-3), "unit": "scores/s",
                              "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12 / world,
                                           "peak": bf16_peak, "unit": "TFLOP/s",
                                           "frac": flops / (ms * 1e-3) / 1e12 / world / bf16_peak}}
    del f, c
    # ---- config 5 ----
    Nall, Dl, k = 1_000_000, 60, 100
    lo, hi = D.shard_bounds(Nall, world, rank)
    ceos = F.normalize(torch.randn(hi - lo, Dl, device=device, generator=g), dim=1)
    firms = F.normalize(torch.randn(hi - lo, Dl, device=device, generator=g), dim=1)
    scale = 1 / 0.07

    def topk_step():
        if world == 1:
            return score_topk(ceos, firms, k, scale)
        return D.score_topk_sharded(ceos, firms, k, scale)

    warm = score_topk(ceos[:4096], firms[:65536], k, scale)
    del warm
    ms = _time_cuda(topk_step, 2, dist, device)      # best of 2: the first full-size call also pays the cudaMallocs
    flops = 2.0 * Nall * Nall * Dl
    out["allpairs_top100"] = {"workload": "config5: 1M CEOs x 1M firms, D=60, top-100 per CEO, exact fp64 rescoring",
                              "ms": ms, "value": float(Nall) * Nall / (ms * 1e-3), "unit": "scores/s",
                              "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12 / world,
                                           "peak": bf16_peak, "unit": "TFLOP/s",
                                           "frac": flops / (ms * 1e-3) / 1e12 / world / bf16_peak}}
    if world == 1:
        # ---- retrieval ranks: rank of the true partner of every row among all columns, counted in the
        # epilogue of the same tensor-core kernel; positives correlated with their rows as after training ----
        from ceo_firm_matching.scoring import diagonal_ranks
        Nr = 262_144
        fr = firms[:Nr]
        cr = F.normalize(0.8 * fr + F.normalize(torch.randn(Nr, Dl, device=device, generator=g), dim=1), dim=1)
        diagonal_ranks(fr[:8192], cr[:8192], method="tensor")
        ms_t = _time_cuda(lambda: diagonal_ranks(fr, cr, method="tensor"), 2, None, device)
        Ne = 16_384
        ms_e = _time_cuda(lambda: diagonal_ranks(fr[:Ne], cr, method="exact"), 1, None, device)
        out["retrieval_ranks"] = {"workload": "rank of the diagonal, 262144 firms x 262144 CEOs, D=60 (contrastive.py:296-332 "
                                              "without its 5000-row cap), fp16 tensor-core filter + exact fp64 near-ties",
                                  "ms": ms_t, "value": float(Nr) * Nr / (ms_t * 1e-3), "unit": "scores/s",
                                  "exact_fp64_simt_kernel_ms_scaled": ms_e * (Nr / Ne),
                                  "roofline": {"bound": "tensor", "achieved": 2.0 * Nr * Nr * Dl / (ms_t * 1e-3) / 1e12,
                                               "peak": bf16_peak, "unit": "TFLOP/s",
                                               "frac": 2.0 * Nr * Nr * Dl / (ms_t * 1e-3) / 1e12 / bf16_peak}}
    return out


def _zipf_indices(n, rows, s_exp, device, gen):
    """Zipf(s) over [0, rows) by inverse-CDF sampling (rank r has weight (r + 1)^-s), SURVEY 8(d) config 4(b)."""
    w = torch.arange(1, rows + 1, dtype=torch.float64, device=device).pow(-s_exp)
    cdf = torch.cumsum(w, 0)
    u = torch.rand(n, dtype=torch.float64, device=device, generator=gen) * cdf[-1]
    return torch.searchsorted(cdf, u).clamp_(max=rows - 1)


def _graph_ms(model, batches, steps, stream=None):
    """ms per step of the graph-captured train step (fwd + loss + bwd) over `batches`."""
    from ceo_firm_matching.training import GraphedTwoTowerStep
    runner = GraphedTwoTowerStep(model, batches[0], optimizer=None, warmup=3, stream=stream)
    for i in range(3):
        runner.step(batches[i % len(batches)])
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        runner.step(batches[i % len(batches)])
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps, runner


def index_distribution_legs(device, batches, steps, uniform_ms):
    """Config 4 under the other index distributions SURVEY 8(d) names: Zipf(1.05) over the 1M-row tables, and the
    reference's real cardinalities (2-4 classes per table, data.py:120-126) at the same batch size: both stress the
    embedding-gradient reduce with long runs of equal keys."""
    from ceo_firm_matching import CEOFirmMatcher, Config
    out = {}
    g = torch.Generator(device=device).manual_seed(77)
    B = B_PER_GPU
    zb = []
    for b in batches[:4]:
        f_cat = torch.stack([_zipf_indices(B, TABLE_ROWS, 1.05, device, g) for _ in range(4)], 1)
        c_cat = torch.stack([_zipf_indices(B, TABLE_ROWS, 1.05, device, g) for _ in range(7)], 1)
        zb.append((b[0], f_cat, b[2], c_cat, b[4], b[5]))
    model = build_model(device, "fp32")
    ms, runner = _graph_ms(model, zb, steps)
    out["config4_zipf"] = {"workload": "config4 with Zipf(s=1.05) categorical indices over the 1M-row tables",
                           "ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "pairs/s",
                           "vs_uniform_indices": ms / uniform_ms}
    del runner, model
    torch.cuda.empty_cache()
    f_cards, c_cards = [4, 4, 2, 2], [2, 4, 2, 2, 2, 2, 2]
    torch.manual_seed(0)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    model = CEOFirmMatcher(meta, Config()).to(device).train()
    model.use_persistent_table_grads(True)
    rb = []
    for b in batches[:4]:
        f_cat = torch.stack([torch.randint(0, n, (B,), device=device, generator=g) for n in f_cards], 1)
        c_cat = torch.stack([torch.randint(0, n, (B,), device=device, generator=g) for n in c_cards], 1)
        rb.append((b[0], f_cat, b[2], c_cat, b[4], b[5]))
    ms, runner = _graph_ms(model, rb, steps)
    out["config4_reference_cardinalities"] = {
        "workload": "config4 batch shape with the reference's table sizes [4,4,2,2] / [2,4,2,2,2,2,2] (runs of 16k-32k equal keys)",
        "ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "pairs/s", "vs_uniform_indices": ms / uniform_ms}
    del runner, model
    torch.cuda.empty_cache()
    return out


def small_config_legs(device, with_cpu):
    """BASELINE configs 1 and 2 (the reference's own CLI runs): device time of one train step and wall-clock of the
    whole training loop, beside a CPU loop of the same structure (the product's CPU data pipeline, which is the
    reference's, + the oracle's torch-CPU step + torch.optim.Adam) on the host cores."""
    import contextlib
    import io
    from sklearn.model_selection import train_test_split
    from torch.utils.data import DataLoader
    import oracle
    from ceo_firm_matching import Config, StructuralConfig
    from ceo_firm_matching.data import CEOFirmDataset, DataProcessor
    from ceo_firm_matching.synthetic import generate_synthetic_data
    from ceo_firm_matching.training import train_model, GraphedTwoTowerStep, BATCH_KEYS
    from ceo_firm_matching.optim import FusedAdam
    from ceo_firm_matching import CEOFirmMatcher
    out = {}
    quiet = contextlib.redirect_stdout(io.StringIO())
    # ---------------- config 1: python -m ceo_firm_matching.cli --synthetic (cli.py:29-59) ----------------
    cfg = Config()
    proc = DataProcessor(cfg)
    with quiet:
        df = proc.prepare_features(generate_synthetic_data(1000))
    train_df, val_df = train_test_split(df, test_size=0.2, random_state=42)
    with quiet:
        proc.fit(train_df)
        train_data, val_data = proc.transform(train_df), proc.transform(val_df)
    train_loader = DataLoader(CEOFirmDataset(train_data), batch_size=256, shuffle=True)
    val_loader = DataLoader(CEOFirmDataset(val_data), batch_size=256, shuffle=False)
    with quiet:
        train_model(train_loader, val_loader, train_data, cfg)              # warm-up: library load, allocator, graphs
    loop_s = float("inf")
    for _ in range(3):                                  # a host-bound loop of ~0.1 s: best of 3 guards against hiccups
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with quiet:
            model = train_model(train_loader, val_loader, train_data, cfg)
        torch.cuda.synchronize()
        loop_s = min(loop_s, time.perf_counter() - t0)
    # one step (fwd + loss + bwd + Adam) at B = 256 as a graph replay
    batch = [train_data[k][:256].to(device) for k in BATCH_KEYS]
    model.train()
    model.use_persistent_table_grads(True)
    opt = FusedAdam(model.parameters(), lr=cfg.LEARNING_RATE)
    st = torch.cuda.Stream(device)
    with torch.cuda.stream(st):
        from ceo_firm_matching.training import eager_step
        eager_step(model, opt, batch)
        torch.cuda.synchronize()
    runner = GraphedTwoTowerStep(model, batch, optimizer=opt, warmup=1, stream=st)
    for _ in range(5):
        runner.step(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        runner.step(batch)
    e1.record()
    torch.cuda.synchronize()
    step_us = e0.elapsed_time(e1) / 200 * 1e3
    steps_per_loop = cfg.EPOCHS * len(train_loader)
    out["config1_cli_synthetic"] = {
        "workload": "config1: cli --synthetic (800 train rows, batch 256, %d epochs, Adam), tables of 2-4 classes" % cfg.EPOCHS,
        "step_us": step_us, "value": 256 / (step_us * 1e-6), "unit": "pairs/s",
        "train_loop_seconds": loop_s, "train_loop_pairs_per_sec": cfg.EPOCHS * len(train_df) / loop_s,
        "steps_per_loop": steps_per_loop}
    del runner
    if with_cpu:
        torch.set_num_threads(os.cpu_count())
        p = oracle.init_two_tower_params(train_data["n_firm_numeric"], train_data["firm_cat_counts"],
                                         train_data["n_ceo_numeric"], train_data["ceo_cat_counts"], seed=0)
        names = [k for k, v in p.items() if v.is_floating_point() and "running" not in k]
        for k in names:
            p[k].requires_grad_(True)
        copt = torch.optim.Adam([p[k] for k in names], lr=cfg.LEARNING_RATE)
        t0 = time.perf_counter()
        n_steps = 0
        for epoch in range(cfg.EPOCHS):
            for b in train_loader:                                   # the reference's per-sample Dataset + collate path
                copt.zero_grad()
                preds = oracle.two_tower_forward(p, b["firm_numeric"], b["firm_cat"], b["ceo_numeric"], b["ceo_cat"],
                                                 training=True)
                oracle.weighted_mse(preds, b["target"], b["weights"]).backward()
                copt.step()
                n_steps += 1
        cpu_s = time.perf_counter() - t0
        out["config1_cli_synthetic"]["cpu_loop"] = {
            "train_loop_seconds": cpu_s, "step_us": cpu_s / n_steps * 1e6, "cores": os.cpu_count(), "kind": "port",
            "sample": "the whole loop: %d steps (DataLoader over CEOFirmDataset + oracle step + torch Adam, dropout off)" % n_steps}
    # ---------------- config 2: structural_cli --synthetic --epochs 100 --batch-size 128 ----------------
    from ceo_firm_matching.structural_data import StructuralDataProcessor
    from ceo_firm_matching.structural_training import train_structural_model
    scfg = StructuralConfig()
    scfg.EPOCHS, scfg.BATCH_SIZE, scfg.DATA_PATH = 100, 128, "SYNTHETIC_MODE"
    sproc = StructuralDataProcessor(scfg)
    with quiet:
        train_ds, val_ds, _ = sproc.load_and_prep()
    s_train = DataLoader(train_ds, batch_size=128, shuffle=True, drop_last=True)
    s_val = DataLoader(val_ds, batch_size=128, shuffle=False)
    warm = StructuralConfig()
    warm.EPOCHS, warm.BATCH_SIZE, warm.DATA_PATH = 2, 128, "SYNTHETIC_MODE"
    with quiet:
        train_structural_model(s_train, s_val, sproc.get_metadata(), warm)
    s_loop = float("inf")
    for _ in range(2):                                  # best of 2 (see config 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with quiet:
            smodel = train_structural_model(s_train, s_val, sproc.get_metadata(), scfg)
        torch.cuda.synchronize()
        s_loop = min(s_loop, time.perf_counter() - t0)
    n_train_steps = scfg.EPOCHS * len(s_train)
    out["config2_structural_cli"] = {
        "workload": "config2: structural_cli --synthetic --epochs 100 --batch-size 128 (1600 train / 400 val rows, KL distillation, Adam)",
        "train_loop_seconds": s_loop, "train_steps": n_train_steps,
        "step_us_incl_validation_share": s_loop / n_train_steps * 1e6,
        "value": scfg.EPOCHS * len(train_ds) / s_loop, "unit": "train pairs/s (whole loop incl. the validation pass)"}
    del smodel
    if with_cpu:
        meta = sproc.get_metadata()
        ps = oracle.init_structural_params(meta["n_firm_num"], meta["firm_cat_cards"], meta["n_ceo_num"], meta["ceo_cat_cards"], seed=0)
        names = [k for k, v in ps.items() if v.is_floating_point() and "running" not in k and k != "A"]
        for k in names:
            ps[k].requires_grad_(True)
        copt = torch.optim.Adam([ps[k] for k in names], lr=scfg.LEARNING_RATE)
        t0 = time.perf_counter()
        n_steps = 0
        for epoch in range(20):                                      # 20 of the 100 epochs: bounded sample
            for b in s_train:
                copt.zero_grad()
                cl, fl, _ = oracle.structural_forward(ps, b["firm_num"], b["firm_cat"], b["ceo_num"], b["ceo_cat"], training=True)
                oracle.structural_kl_loss(cl, fl, b["target_ceo"], b["target_firm"]).backward()
                copt.step()
                n_steps += 1
            with torch.no_grad():
                for b in s_val:
                    cl, fl, _ = oracle.structural_forward(ps, b["firm_num"], b["firm_cat"], b["ceo_num"], b["ceo_cat"], training=False)
                    oracle.structural_kl_loss(cl, fl, b["target_ceo"], b["target_firm"])
        cpu_s = (time.perf_counter() - t0) * (scfg.EPOCHS / 20)
        out["config2_structural_cli"]["cpu_loop"] = {
            "train_loop_seconds": cpu_s, "cores": os.cpu_count(), "kind": "port",
            "sample": "20 of the 100 epochs timed (DataLoader + oracle step + torch Adam + validation pass), scaled x5"}
    return out


def gpu_arm(args):
    from ceo_firm_matching import _native as N
    import ctypes as C
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    lib = N.lib()
    model = build_model(device, args.precision)
    dp = None
    if world > 1:
        from ceo_firm_matching import distributed as D
        # default: towers data-parallel, tables sharded over NVLink peer memory (per-rank exchange independent of
        # the world size); --tables replicated = the all-gather form SURVEY 8e describes (does not scale)
        dp = (D.TableShardedTwoTower(model, batch_rows=B_PER_GPU) if args.tables == "sharded"
              else D.DataParallelTwoTower(model))

    n_data = max(8, min(args.steps + args.warmup, N_PAIRS // B_PER_GPU // max(world, 1)))
    batches = make_batches(n_data, B_PER_GPU, device, seed=1234 + rank)     # ~10 MB each: >> L2 in total

    from ceo_firm_matching.training import GraphedTwoTowerStep
    # the whole step (sparse re-zero, fwd, loss, bwd, segment reduce) is captured once and replayed
    # multi-GPU: the gradient synchronisation (NCCL all-reduce + the owners' peer reduce) is captured with the step
    runner = GraphedTwoTowerStep(model, batches[0], optimizer=None, warmup=3,
                                 loss_scale=dp.loss_scale if dp is not None else 1.0,
                                 after_backward=dp.sync_gradients if dp is not None else None,
                                 before_forward=getattr(dp, "begin_step", None))

    def step(i):
        return runner.step(batches[i % n_data])      # D2D copy into the graph's static inputs + replay

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        for i in range(max(args.warmup, 3)):
            step(i)
        barrier()
        clocks.mark()
        e0.record()
        for i in range(args.steps):
            loss = step(args.warmup + i)
        e1.record()
        barrier()
        clocks.unmark()
    ms_total = e0.elapsed_time(e1)

    # Per-kernel durations: a graph replay hides the individual launches from host-side events, so the same
    # steps are re-issued eagerly (identical kernels, identical arguments) with a cudaEvent pair around every
    # kernel family on the launching stream.  Launch count: kernels of libcfm_b200 per step x timed steps.
    from ceo_firm_matching.training import eager_step
    with torch.cuda.stream(runner.stream):
        eager_step(model, None, batches[0])
        torch.cuda.synchronize()
        lib.cfm_launch_count(1)
        lib.cfm_profile_enable(1)
        n_prof = min(args.steps, 10)
        for i in range(n_prof):
            eager_step(model, None, batches[i % n_data])
        torch.cuda.synchronize()
    launches = int(lib.cfm_launch_count(0)) // n_prof * args.steps
    prof_ms = (C.c_double * 14)()
    prof_n = (C.c_int64 * 14)()
    N.check(lib.cfm_profile_read(prof_ms, prof_n, 14))
    lib.cfm_profile_enable(0)
    model.zero_grad_fast()
    prof_step_ms = sum(prof_ms) / n_prof
    if dist is not None:
        tmax = torch.tensor([ms_total], device=device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / args.steps
    value = world * B_PER_GPU * args.steps / (ms_total * 1e-3)

    # ---- end-to-end: host (pinned) batches, H2D + step + D2H of the loss inside the timed region ----
    host = [[t.cpu().pin_memory() for t in b] for b in batches[:4]]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    copy_stream = torch.cuda.Stream(device)
    loss_host = torch.zeros(1).pin_memory()

    # two preallocated device staging sets (no allocator traffic in the timed loop): the copy stream fills set k while
    # the step consumes set 1-k; events order reuse in both directions
    staging = [[torch.empty(t.shape, dtype=t.dtype, device=device) for t in host[0]] for _ in range(2)]
    consumed = [None, None]

    def upload(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])          # the step that read this set has copied it out
            for d, h in zip(staging[k], host[i % len(host)]):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return k, ev

    def e2e_loop(n):
        nxt = upload(0)
        for i in range(n):
            k, ev = nxt
            if i + 1 < n:
                nxt = upload(i + 1)            # prefetch the next batch while this one computes
            main = torch.cuda.current_stream()
            main.wait_event(ev)
            runner.load(staging[k])            # D2D into the graph's static inputs
            done = torch.cuda.Event()
            done.record(main)
            consumed[k] = done
            loss = runner.step()
            loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.synchronize()

    e2e_loop(3)
    e2e_runs = []
    for _ in range(3):                       # K steps each; the median guards against a one-off host hiccup
        barrier()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            tmax = torch.tensor([e2e_s], device=device)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e2e_s = float(tmax.item())
        e2e_runs.append(e2e_s)
    e2e_s = statistics.median(e2e_runs)
    e2e_value = world * B_PER_GPU * args.steps / e2e_s

    secondary = secondary_metrics(device, world, rank, dist)
    if world == 1 and not args.headline_only:
        secondary.update(index_distribution_legs(device, batches, args.steps, ms_step))
        secondary.update(small_config_legs(device, with_cpu=not args.no_cpu))
    if world == 1:
        # the same step in the other tower-product precision, also as one graph replay per step
        other = "tf32" if args.precision == "fp32" else "fp32"
        model.set_precision(other)
        model.zero_grad_fast()
        runner2 = GraphedTwoTowerStep(model, batches[0], optimizer=None, warmup=3, stream=runner.stream)
        for i in range(3):
            runner2.step(batches[i % n_data])
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            runner2.step(batches[i % n_data])
        t1.record()
        torch.cuda.synchronize()
        ms_o = t0.elapsed_time(t1) / args.steps
        secondary["two_tower_" + other] = {"workload": WORKLOADS[other], "ms_per_step": ms_o,
                                           "value": B_PER_GPU / (ms_o * 1e-3), "unit": "pairs/s"}
        model.set_precision(args.precision)
        model.zero_grad_fast()

    if world == 1:
        # ---- the optimiser the metric excludes: the same step followed by Adam over all 1.4 GB of
        # parameters, torch.optim.Adam(capturable, foreach) vs the fused single-pass kernel (bit-equal results) ----
        from ceo_firm_matching.optim import FusedAdam
        n_param = sum(p.numel() for p in model.parameters())
        adam = {}
        for name, make in (("fused", lambda: FusedAdam(model.parameters(), lr=1e-3)),
                           ("torch_foreach", lambda: torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True))):
            opt = make()
            model.zero_grad_fast()
            with torch.cuda.stream(runner.stream):
                eager_step(model, opt, batches[0])               # optimiser state exists before the capture
                torch.cuda.synchronize()
            r3 = GraphedTwoTowerStep(model, batches[0], optimizer=opt, warmup=1, stream=runner.stream)
            for i in range(2):
                r3.step(batches[i % n_data])
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for i in range(10):
                r3.step(batches[i % n_data])
            t1.record()
            torch.cuda.synchronize()
            adam[name] = t0.elapsed_time(t1) / 10
            del r3, opt
            torch.cuda.empty_cache()
        adam_ms = adam["fused"] - ms_step
        secondary["train_step_with_adam"] = {
            "workload": "config4 step + dense Adam over %d parameters (eleven 1M-row tables), one graph replay" % n_param,
            "ms_per_step_fused_adam": adam["fused"], "ms_per_step_torch_adam": adam["torch_foreach"],
            "value": B_PER_GPU / (adam["fused"] * 1e-3), "unit": "pairs/s",
            "roofline": {"bound": "hbm", "kernel": "adam_kernel", "achieved": 28.0 * n_param / (adam_ms * 1e-3) / 1e9,
                         "peak": peaks()[0], "unit": "GB/s", "frac": 28.0 * n_param / (adam_ms * 1e-3) / 1e9 / peaks()[0],
                         "note": "7 floats per parameter (read p, g, m, v; write p, m, v); kernel time = step with Adam - step without"}}
        model.zero_grad_fast()

    if rank != 0:
        _finish(dist, runner)
        return

    # ---- roofline of the dominant kernel (stage-1 backward: dW1 + dX + embedding-row gradients) ----
    peak, peak_src = peaks()
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            ncu_traffic = json.load(f)
    except Exception:
        ncu_traffic = {}
    slots = ["fwd1", "fwd2", "fwd3", "bwd1", "bwd2", "bwd3", "head", "emb_grad", "reduce", "nce_rowsum", "nce_grad",
             "topk", "topk_post", "adam"]
    per_kernel = {s: (prof_ms[i] / prof_n[i] if prof_n[i] else None) for i, s in enumerate(slots)}
    shares = {s: round(prof_ms[i] / n_prof / ms_step, 4) for i, s in enumerate(slots) if prof_n[i]}
    # dominant kernel = the tower stage kernel with the largest share (the embedding-gradient and reduction slots sum
    # several small launches, some of them on the library's side stream)
    dom = max((s for s in slots[:6] if per_kernel[s]), key=lambda s: prof_ms[slots.index(s)])
    dom_bytes = {"bwd1": BYTES_BWD1_PER_PAIR, "fwd1": 1148}.get(dom, BYTES_BWD1_PER_PAIR) * B_PER_GPU
    achieved = dom_bytes / (per_kernel[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch, read from the summary of this
                # round's `ncu --set full` capture at this exact shape (written by scratch/ncu_traffic.py)
                "traffic": ncu_traffic.get(dom), "traffic_unit": "bytes per launch",
                "traffic_source": ncu_traffic.get("source"),
                "algorithmic_bytes_per_launch": dom_bytes,
                "kernel": "tower_%s_tc (tcgen05 kind::tf32, TMEM accumulators)" % ("bwd" if dom.startswith("bwd") else "fwd"),
                "slot": dom, "peak_source": peak_src, "kernel_ms": per_kernel[dom], "kernel_share_of_step": shares,
                "whole_step_frac": BYTES_PER_PAIR * B_PER_GPU / (ms_step * 1e-3) / 1e9 / peak,
                "kernel_sum_ms_per_step": prof_step_ms,
                "note": "per-kernel times from an eager re-issue of the timed steps (graph replays hide them)"}

    cpu_value, cpu_ms = (None, None)
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu_value, cpu_ms = run_cpu(3, 1)
        cpu = {"value": cpu_value, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "3 steps (after 1 warm-up) of one B=65536 batch, fwd+loss+bwd, dense 1M-row table grads"}

    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "tf32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.precision], "global_batch": world * B_PER_GPU,
                   "parallelism": ("single" if world == 1 else
                                   f"dp{world} towers + tables {args.tables}" +
                                   (" over NVLink peer memory" if args.tables == "sharded" else " (all-gather)")),
                   "l2": f"{n_data} distinct 10 MB batches cycled + 992 MB tables (inputs >> 126 MB L2)",
                   "arithmetic": ("tower products as error-compensated 3xTF32 tcgen05 MMAs (fp32-class: parity "
                                  "rtol 2e-5 vs the oracle), everything else fp32" if args.precision == "fp32" else
                                  "tower products single-pass TF32 (fp32 accumulate), everything else fp32"),
                   "optimizer_step": "excluded from the metric (SURVEY 8d)", "dropout": 0.1,
                   "launch": "one CUDA-graph replay per step"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "note": "pinned host batches, copy stream prefetches batch i+1 during step i, loss read back every step; median of 3 timed repetitions of K steps"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "secondary": secondary,
    }
    print(json.dumps(line))
    _finish(dist, runner)


def _finish(dist, runner):
    """Multi-rank exit.  The captured step holds NCCL kernels: the graph is dropped and the ranks rendezvous before the
    process group is destroyed.  destroy_process_group() has been seen to wait forever with such a graph alive, so it
    runs under a watchdog that ends the process (exit code 0, results are already printed) if it does not return."""
    sys.stdout.flush()
    sys.stderr.flush()
    if dist is None:
        return
    runner.graph.reset()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    import threading

    def _bail():
        sys.stderr.write("bench: destroy_process_group() did not return within 15 s; leaving without NCCL teardown\n")
        sys.stderr.flush()
        os._exit(0)

    watchdog = threading.Timer(15.0, _bail)
    watchdog.daemon = True
    watchdog.start()
    dist.destroy_process_group()
    watchdog.cancel()
    os._exit(0)        # peer (CUDA IPC) mappings of the other ranks' tables are still open: skip interpreter teardown


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (debugging)")
    ap.add_argument("--headline-only", action="store_true",
                    help="skip the index-distribution and config 1/2 secondary legs (profiling runs)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"],
                    help="tower products: fp32-class (3xTF32, the reference's precision; default) or single-pass TF32")
    ap.add_argument("--tables", default="sharded", choices=["sharded", "replicated"],
                    help="multi-GPU embedding tables: sharded over NVLink peer memory, or replicated + all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
