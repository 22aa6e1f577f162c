#!/bin/bash
# Final captures of round 2 (after the one-pass InfoNCE forward, the rank mode and the folded BatchNorm finalizes).
# Only CSV exports travel back (gpurun_out is capped at 64 MiB); every program runs once WITHOUT ncu first.
set -u
T=/tmp/ncu; mkdir -p $T gpurun_out
exp() { ncu -i $T/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null; }
python scratch/prof_nce.py > gpurun_out/r02c_p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simtile_kernel -c 2 -o $T/r02c_nce python scratch/prof_nce.py > gpurun_out/r02c_p1n.log 2>&1
exp r02c_nce
python scratch/prof_rank.py > gpurun_out/r02c_p2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simtile_kernel -c 1 -o $T/r02c_rank python scratch/prof_rank.py > gpurun_out/r02c_p2n.log 2>&1
exp r02c_rank
# the six tower stage kernels of the third (warm) step + the launch list of a whole step
python scratch/prof_tower_only.py fp32 > gpurun_out/r02c_p3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tower_ --launch-skip 12 --launch-count 6 -o $T/r02c_towers python scratch/prof_tower_only.py fp32 > gpurun_out/r02c_p3n.log 2>&1
exp r02c_towers
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_step.csv python scratch/prof_tower_only.py fp32 > $T/p3l.log 2>&1
ls -la gpurun_out | tail -8
