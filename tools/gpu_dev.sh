#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/sweep.py --workload c2 --worlds 4096 --steps 20 --variants "fused:1:16" 2>&1 | grep -v "^\["
timeout 300 python tools/sweep.py --workload c3 --worlds 16384 --steps 6 --settle 203 --variants "wide:1:16" 2>&1 | grep -v "^\["
