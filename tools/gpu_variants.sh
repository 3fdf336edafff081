#!/bin/bash
# Times builds of the CUDA library that differ in compile-time knobs on the same settled state.
# Build them first, in-tree so that they travel to the GPU box (tools/variants/ is git-ignored through *.so):
#   nvcc <flags of __graft_entry__.NVCC_FLAGS> -D<KNOB>=<value> -o tools/variants/libps_<X>.so physicssimulator_b200/csrc/ps_capi.cu physicssimulator_b200/csrc/ps_broadphase.cu
# then: gpurun -- 'VARIANTS="X Y" bash tools/gpu_variants.sh'   (the Python binding loads $PS_B200_LIB instead of the product library)
set -u
mkdir -p gpurun_out
run() { timeout 200 python tools/sweep.py --workload $1 --worlds $2 --steps $3 --settle $4 --variants "fused:1:16" 2>&1 | cut -c1-60; }
echo "== product"; run c2 4096 20 210; run c4 8192 10 12
for v in ${VARIANTS:-O}; do echo "== variant $v"; export PS_B200_LIB=$PWD/tools/variants/libps_$v.so; run c2 4096 20 210; run c4 8192 10 12; unset PS_B200_LIB; done
for c in 1 10 25 50 100; do echo "== product carveout $c"; PS_B200_CARVEOUT=$c run c2 4096 20 210; done
PS_B200_LIB=$PWD/tools/variants/libps_O.so timeout 300 python -m pytest tests -m gpu -x -q -k "t7 or c2_c3 or tile_widths" 2>&1 | tail -2
