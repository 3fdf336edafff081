#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "flow" > gpurun_out/pytest_flow.log 2>&1; echo "pytest flow rc=$?"; tail -12 gpurun_out/pytest_flow.log
timeout 200 python tools/sweep.py --workload c2 --worlds 4096 --steps 20 --settle 210 --variants "fused:1:16,flow:1:16,wide:1:16" 2>&1 | grep -v "^\[ps_b200 wide"
timeout 300 python tools/sweep.py --workload c3 --worlds 16384 --steps 6 --settle 210 --variants "wide:1:16,flow:1:16" 2>&1 | grep -v "^\[ps_b200 wide"
