#!/bin/bash
set -u
mkdir -p gpurun_out
for v in mb2 mb3; do
  cp tools/micro/libs/$v.so physicssimulator_b200/libps_b200.so
  echo "== $v"
  timeout 300 python tools/sweep.py --workload c3 --worlds 16384 --steps 8 --settle 210 --variants "wide:1:16,wide:1:16" 2>&1 | grep -v "^\["
done
timeout 300 python tools/sweep.py --workload c3 --worlds 16384 --steps 4 --settle 210 --prof --variants "wide:1:16" > gpurun_out/rounds_c3_mb3.txt 2>&1; grep "wide profile\] steps" gpurun_out/rounds_c3_mb3.txt; grep "round   1:\|round  10:\|round  20:" gpurun_out/rounds_c3_mb3.txt
