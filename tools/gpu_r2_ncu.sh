#!/bin/bash
# ncu --set full of one mid-round launch of the resolve and box-box kernels of a settled C3 step (bench --profile-step)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-extra --rollout 0 --profile-step"
export PS_B200_TEAM_MAX=0
for k in ${KERNELS:-wide_round_resolve_d wide_round_bb_d}; do
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -s ${SKIP:-4} -c 1 -o gpurun_out/r2_ncu_$k -f $CMD > gpurun_out/r2_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
