"""bench.py's CPU legs (the reference arm / cpu_baseline) at a shrunken size: they must run without a GPU and print
the contract's JSON keys."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "B_PER_GPU", 256)
    monkeypatch.setattr(bench, "TABLE_ROWS", 1000)
    monkeypatch.setattr(bench, "F_CARDS", [1000] * 4)
    monkeypatch.setattr(bench, "C_CARDS", [1000] * 7)
    monkeypatch.setenv("WORLD_SIZE", "1")
    monkeypatch.setenv("RANK", "0")

    class Args:
        steps, warmup, gpus, precision = 1, 1, 1, "fp32"

    bench.reference_arm(Args())
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "pairs/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == bench.WORKLOADS["fp32"]


def test_clock_sampler_degrades_without_a_gpu():
    """On a box with neither NVML nor nvidia-smi the sampler must not raise; it reports zero samples."""
    sys.path.insert(0, ROOT)
    import bench
    with bench.ClockSampler(0) as clocks:
        clocks.mark()
        clocks.unmark()
    summary = clocks.summary()
    assert set(summary) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source", "window"}
    assert summary["samples"] >= 0 and isinstance(summary["reasons"], list)
