#!/bin/bash
# Round 2 iteration session: parity tests, then timings of the batch-wide path from the same settled state.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest_gpu.log
run() { timeout 300 python tools/sweep.py --workload $1 --worlds $2 --steps $3 --settle $4 --variants "$5" 2>&1 | cut -c1-200; }
echo "== c3 16384"; run c3 16384 10 210 "wide:1:16"
for tm in ${TEAMS:-0 1000000}; do echo "== c3 team_max=$tm"; PS_B200_TEAM_MAX=$tm run c3 16384 10 210 "wide:1:16"; done
echo "== c3 host rounds"; PS_B200_WIDE_HOST_ROUNDS=1 run c3 16384 10 210 "wide:1:16"
echo "== c2 4096"; run c2 4096 20 210 "wide:1:16,fused:1:16"
echo "== c3 2048"; run c3 2048 10 210 "wide:1:16,fused:1:16"
