"""Repeat bench.small_config_legs (config 1 / config 2 whole-loop seconds) to see the spread between runs."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ceo-recommender_b200')
import bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
for i in range(3):
    out = bench.small_config_legs(dev, with_cpu=False)
    print(i, out["config1_cli_synthetic"]["train_loop_seconds"], out["config1_cli_synthetic"]["step_us"],
          out["config2_structural_cli"]["train_loop_seconds"], flush=True)
