#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest_gpu.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-parity --no-extra --rollout 0"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'e2e', round(d['e2e']['ms_per_step'],2), 'fallbacks', d['stats']['fallback_steps'])"; }
PS_B200_WIDE_HOST_ROUNDS=1 timeout 600 $B 2>/dev/null | pick "host-rounds"
timeout 600 $B 2>/dev/null | pick "device-counted"
PS_B200_TEAM_MAX=0 timeout 600 $B 2>/dev/null | pick "device-counted-noteam"
PS_B200_TEAM_MAX=100000 timeout 600 $B 2>/dev/null | pick "device-counted-team100k"
timeout 600 $B --workload c2 2>/dev/null | pick "c2 fused"
PS_B200_PATH=wide timeout 600 $B --workload c2 2>/dev/null | pick "c2 wide"
timeout 600 python tools/sweep.py --workload c3 --worlds 16384 --steps 6 --settle 210 --variants "wide:1:16" --prof > gpurun_out/r2e_wide_rounds_c3.txt 2>&1; head -8 gpurun_out/r2e_wide_rounds_c3.txt | cut -c1-200
