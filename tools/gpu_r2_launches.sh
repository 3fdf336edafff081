#!/bin/bash
# launch list (durations + DRAM bytes) of ONE settled step of the bench workload
set -u
mkdir -p gpurun_out
W=${W:-c3}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-cpu --no-parity --no-extra --rollout 0 --profile-step"
timeout 900 ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_launches_${W}_step.csv $CMD > gpurun_out/r2_ncu_${W}_step.log 2>&1; echo "ncu $W step rc=$?"
