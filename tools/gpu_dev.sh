#!/bin/bash
# One short GPU session while iterating: parity tests, C2/C3/C4 timing from the same settled state, optional full ncu capture.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { timeout 200 python tools/sweep.py --workload $1 --worlds $2 --steps $3 --settle $4 --variants "fused:1:16" 2>&1 | cut -c1-60; }
echo "== re-ordered"; export PS_B200_PERM=1; run c2 4096 40 210; run c4 8192 20 12; run c2 1024 40 210
echo "== creation order"; export PS_B200_PERM=0; run c2 4096 40 210; run c4 8192 20 12; run c2 1024 40 210
export PS_B200_PERM=1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench c2', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['sequential_calls_value_this_rank'])"
if [ "${NCU:-0}" = "1" ]; then
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --settle 100"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/prof_c2_dev -f $CMD > gpurun_out/ncu_dev.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dev.log
fi
