#!/bin/bash
# One short GPU session while iterating: parity tests, C2/C4 timing of the product against a variant build from the same settled state.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { timeout 200 python tools/sweep.py --workload $1 --worlds $2 --steps $3 --settle $4 --variants "fused:1:16" 2>&1 | cut -c1-60; }
echo "== product"; run c2 4096 40 210; run c4 8192 20 12; run c2 1024 40 210; run c3 2048 10 210
if [ -f tools/variants/libps_L.so ]; then
echo "== variant L"; export PS_B200_LIB=$PWD/tools/variants/libps_L.so; run c2 4096 40 210; run c4 8192 20 12; run c2 1024 40 210; run c3 2048 10 210; unset PS_B200_LIB
fi
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench c2', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['sequential_calls_value_this_rank'])"
timeout 200 python tools/stress_fast_vs_serial.py --workload c2 --worlds 128 --steps 300 --path fused 2>&1 | tail -1 | cut -c1-200
timeout 200 python tools/stress_fast_vs_serial.py --workload c3 --worlds 64 --steps 150 --path fused 2>&1 | tail -1 | cut -c1-200
