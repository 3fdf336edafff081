#!/bin/bash
set -u
mkdir -p gpurun_out
for v in mb16 mb14; do
  cp tools/micro/libs/$v.so physicssimulator_b200/libps_b200.so
  echo "== $v"
  timeout 300 python tools/sweep.py --workload c2 --worlds 4096 --steps 20 --settle 210 --variants "fused:1:16,fused:1:16" 2>&1 | grep -v "^\["
  timeout 300 python tools/sweep.py --workload c4 --worlds 8192 --steps 10 --settle 20 --variants "fused:1:16" 2>&1 | grep -v "^\["
done
