#!/bin/bash
# Round-2 ncu captures; only CSV exports travel back (reports stay on the box: gpurun_out is capped at 64 MiB).
set -u
T=/tmp/ncu; mkdir -p $T gpurun_out
exp() {  # report -> raw csv (+ optional source csv of launch $3)
  ncu -i $T/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  if [ -n "${2:-}" ]; then ncu -i $T/$1.ncu-rep --page source --csv --print-source sass --launch-skip $2 --launch-count 1 > gpurun_out/$1.src$2.csv 2>/dev/null; fi
}
python scratch/prof_nce.py > $T/p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simtile_kernel -c 3 -o $T/r02_nce_full2 python scratch/prof_nce.py > gpurun_out/p1n.log 2>&1
exp r02_nce_full2 2
python scratch/prof_tower_only.py fp32 > $T/p2.log 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 90 --launch-count 45 -o $T/r02_step_full2 python scratch/prof_tower_only.py fp32 > gpurun_out/p2n.log 2>&1
exp r02_step_full2
# index of the stage-1 backward inside the capture, for its source page
IDX=$(python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r02_step_full2.raw.csv")))
kn=rows[0].index("Kernel Name")
for i,r in enumerate(rows[2:]):
    if "tower_bwd_tc<(bool)1" in r[kn]: print(i); break
PY
)
[ -n "$IDX" ] && ncu -i $T/r02_step_full2.ncu-rep --page source --csv --print-source sass --launch-skip $IDX --launch-count 1 > gpurun_out/r02_step_full2.bwd1.src.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_step2.csv python scratch/prof_tower_only.py fp32 > $T/p2l.log 2>&1
python scratch/prof_topk_cfg5.py > $T/p3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simtile_kernel -c 1 -o $T/r02_topk_cfg5 python scratch/prof_topk_cfg5.py > gpurun_out/p3n.log 2>&1
exp r02_topk_cfg5 0
ls -la gpurun_out | tail -12
