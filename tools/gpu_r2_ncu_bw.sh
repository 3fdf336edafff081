#!/bin/bash
# ncu --set full of the bandwidth-bound kernels: wide_predict / wide_final / wide_near_list (one settled C3 step) and the
# broad-phase kernels; the reports are summarised on the box (tools/ncu_hot.py) and only the summaries come back
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-extra --rollout 0 --profile-step"
for k in wide_predict wide_final wide_near_list wide_round_resolve_d wide_round_bb_d wide_round_detect_d; do
  S=0; case $k in wide_round_*) S=9;; esac
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:$k -s $S -c 1 -o /tmp/r2_ncu_$k -f $CMD > gpurun_out/r2_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_hot.py /tmp/r2_ncu_$k.ncu-rep 24 > gpurun_out/r2_ncu_$k.txt 2>&1
  ncu -i /tmp/r2_ncu_$k.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$k.raw.csv 2>/dev/null
done
for k in own_pairs_kernel scatter_kernel aabb_kernel; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -o /tmp/r2_ncu_bp_$k -f python tools/bench_broadphase.py 1 > gpurun_out/r2_ncu_bp_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/ncu_hot.py /tmp/r2_ncu_bp_$k.ncu-rep 12 > gpurun_out/r2_ncu_bp_$k.txt 2>&1
  ncu -i /tmp/r2_ncu_bp_$k.ncu-rep --page raw --csv > gpurun_out/r2_ncu_bp_$k.raw.csv 2>/dev/null
done
du -sh gpurun_out
