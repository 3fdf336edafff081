#!/bin/bash
# Times builds of the CUDA library that differ in compile-time knobs (tools/variants/libps_<X>.so) on the same settled state.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-A B C D E}; do
  for t in ${TILES:-16}; do
    echo "== variant $v tile $t"
    PS_B200_LIB=$PWD/tools/variants/libps_$v.so timeout 200 python tools/sweep.py --workload c2 --worlds 4096 --steps 20 --settle 210 --variants "fused:1:16:$t" 2>&1 | grep -v "^\[ps_b200 wide" | cut -c1-120
  done
done
if [ -n "${EXTRA:-}" ]; then bash -c "$EXTRA"; fi
