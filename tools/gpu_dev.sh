#!/bin/bash
# One short GPU session while iterating: parity tests, C2/C3/C4 timing from the same settled state, one full ncu capture.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/sweep.py --workload c2 --worlds 4096 --steps 20 --settle 210 --variants "fused:1:16,fused:1:16:16,wide:1:16" 2>&1 | grep -v "^\[ps_b200 wide" | cut -c1-150
timeout 200 python tools/sweep.py --workload c4 --worlds 8192 --steps 10 --settle 12 --variants "fused:1:16" 2>&1 | cut -c1-150
timeout 300 python tools/sweep.py --workload c3 --worlds 16384 --steps 6 --settle 210 --variants "wide:1:16" 2>&1 | grep -v "^\[ps_b200 wide" | cut -c1-150
timeout 200 python tools/bench_c1.py 2>&1 | tail -1 | cut -c1-200
if [ "${NCU:-1}" = "1" ]; then
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --settle 100"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/prof_c2_dev -f $CMD > gpurun_out/ncu_dev.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dev.log
fi
