"""Per-launch DRAM traffic and headline metrics of the step's kernels from an `ncu --set full` report.

usage: python scratch/ncu_traffic.py gpurun_out/r02_step_full2.raw.csv [more .raw.csv / .ncu-rep ...]
writes profiles/r02_ncu_traffic.json (read by bench.py for `roofline.traffic`) and prints a markdown table
(kernel, us, dram read / write MB, warps active, issue active, tensor pipe, registers) for profiles/r02_ncu_summary.md."""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "us": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lsu_smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "smem_kb": "launch__shared_mem_per_block_dynamic",
}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0,
        "nsecond": 1e-3}


def slot_of(name):
    m = re.search(r"tower_fwd_tc<\(int\)(\d+), \(bool\)(\d)>", name)
    if m:
        return "fwd1" if m.group(2) == "1" else None
    m = re.search(r"tower_bwd_tc<\(bool\)(\d), \(int\)(\d+)>", name)
    if m:
        return "bwd1" if m.group(1) == "1" else None
    return None


def rows_of(path):
    if path.endswith(".csv"):                        # already exported with `ncu -i x.ncu-rep --page raw --csv`
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(v) for k, v in WANT.items() if v in hdr}
    kn = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        d = {"name": r[kn]}
        for k, i in idx.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                v = None
            if v is not None and k in ("us", "rd", "wr"):
                v *= UNIT.get(units[i], 1.0)
            d[k] = v
        res.append(d)
    return res


def main():
    traffic = {"source": "ncu --set full --clock-control none, " + ", ".join(os.path.basename(p) for p in sys.argv[1:]) +
                         " (profiles/r02_ncu_summary.md); dram__bytes_read.sum + dram__bytes_write.sum per launch"}
    print("| kernel | us | dram rd MB | dram wr MB | warps active % | issue active % | tensor pipe % | XU % | FMA % | LSU smem % | regs |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    stage_names = {}
    for path in sys.argv[1:]:
        for d in rows_of(path):
            short = re.sub(r"\(.*$", "", d["name"]).replace("void ", "").replace("cfm::", "")
            short = re.sub(r"\(int\)|\(bool\)", "", short)
            label = short
            m = re.search(r"tower_(fwd|bwd)_tc<(\d+), (\d+)>", short)
            if m and m.group(1) == "fwd":          # <accumulator columns, stage 1?>: config 4 has N = 64, 32, 60(->64)
                st = 1 if m.group(3) == "1" else (2 if m.group(2) == "32" else 3)
                label += f" (fwd stage {st})"; slot = f"fwd{st}"
            elif m:                                # <stage 1?, input-column chunk width>: K = 204/58, 64, 32
                st = 1 if m.group(2) == "1" else (2 if m.group(3) == "64" else 3)
                label += f" (bwd stage {st})"; slot = f"bwd{st}"
            else:
                slot = None
            f = lambda k, n=1: "-" if d.get(k) is None else f"{d[k]:.{n}f}"
            rd, wr = (d.get("rd") or 0) / 1e6, (d.get("wr") or 0) / 1e6
            print(f"| `{label[:70]}` | {f('us')} | {rd:.1f} | {wr:.1f} | {f('warps')} | {f('issue')} | {f('tensor')} | {f('xu')} | {f('fma')} | {f('lsu_smem')} | {f('regs', 0)} |")
            if slot and d.get("rd") is not None:
                traffic[slot] = int((d["rd"] or 0) + (d["wr"] or 0))
                traffic[slot + "_us_under_ncu"] = d.get("us")
    with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main()
