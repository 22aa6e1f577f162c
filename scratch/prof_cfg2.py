"""Where the config-2 loop (structural_cli --synthetic --epochs 100 --batch-size 128) spends its wall time."""
import sys, time, io, contextlib, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
from torch.utils.data import DataLoader
from ceo_firm_matching import StructuralConfig
from ceo_firm_matching.structural_data import StructuralDataProcessor
from ceo_firm_matching import structural_training as ST
from ceo_firm_matching.batching import device_batches
from ceo_firm_matching import ops
quiet = contextlib.redirect_stdout(io.StringIO())
scfg = StructuralConfig(); scfg.EPOCHS, scfg.BATCH_SIZE, scfg.DATA_PATH = 100, 128, "SYNTHETIC_MODE"
sproc = StructuralDataProcessor(scfg)
with quiet:
    train_ds, val_ds, _ = sproc.load_and_prep()
s_train = DataLoader(train_ds, batch_size=128, shuffle=True, drop_last=True)
s_val = DataLoader(val_ds, batch_size=128, shuffle=False)
warm = StructuralConfig(); warm.EPOCHS, warm.BATCH_SIZE, warm.DATA_PATH = 2, 128, "SYNTHETIC_MODE"
with quiet:
    ST.train_structural_model(s_train, s_val, sproc.get_metadata(), warm)
torch.cuda.synchronize()
t0 = time.perf_counter()
with quiet:
    model = ST.train_structural_model(s_train, s_val, sproc.get_metadata(), scfg)
torch.cuda.synchronize()
print("whole loop s", time.perf_counter() - t0, "train steps/epoch", len(s_train), "val steps", len(s_val))
# pieces
dev = torch.device("cuda")
from ceo_firm_matching.optim import FusedAdam
opt = FusedAdam(model.parameters(), lr=1e-3)
batches = list(device_batches(s_train, ST.BATCH_KEYS, dev))
stream = torch.cuda.Stream(dev)
with torch.cuda.stream(stream):
    model.train()
    opt.zero_grad(set_to_none=True); l = ST._loss(model, batches[0]); l.backward(); opt.step()
    g = ST.GraphedStructuralStep(model, opt, batches[0], stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(200): g.step(batches[i % len(batches)])
    e1.record(); torch.cuda.synchronize()
    print("graph step: gpu us", e0.elapsed_time(e1) * 5, "wall us", (time.perf_counter() - t0) / 200 * 1e6)
model.eval()
vb = list(device_batches(s_val, ST.BATCH_KEYS, dev))
with torch.no_grad():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(100): ST._loss(model, vb[i % len(vb)])
    torch.cuda.synchronize(); print("eager val step wall us", (time.perf_counter() - t0) / 100 * 1e6)
t0 = time.perf_counter()
for _ in range(20):
    for b in device_batches(s_train, ST.BATCH_KEYS, dev): pass
torch.cuda.synchronize(); print("device_batches per epoch us", (time.perf_counter() - t0) / 20 * 1e6)
