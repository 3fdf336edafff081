#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py --workload c2 --worlds 4096 --steps 20 --variants "fused:2:16,fused:1:16,fused:1:4,fused:0.7:16,wide:2:16,wide:1:16,wide:0.7:16" 2>&1 | grep -v "^\[" | tee gpurun_out/sweep_c2.txt
timeout 300 python tools/sweep.py --workload c2 --worlds 4096 --steps 8 --prof --variants "wide:1:16" 2>&1 | tee gpurun_out/sweep_c2_prof.txt | tail -3
timeout 300 python tools/sweep.py --workload c3 --worlds 16384 --steps 8 --settle 203 --prof --variants "wide:1:16" 2>&1 | tee gpurun_out/sweep_c3_prof.txt | tail -3
