#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python tools/sweep.py --workload c3 --worlds 16384 --steps 2 --settle 203 --variants wide:1:16"
ncu --set full --clock-control none --import-source on -k regex:wide_round_bb -s 8600 -c 1 -o gpurun_out/prof_c3_bb -f $CMD > gpurun_out/ncu_bb.log 2>&1; echo "ncu bb rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wide_round_resolve -s 8600 -c 1 -o gpurun_out/prof_c3_res -f $CMD > gpurun_out/ncu_res.log 2>&1; echo "ncu res rc=$?"
