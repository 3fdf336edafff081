#!/bin/bash
set -u
run() { timeout 300 python tools/sweep.py --workload $1 --worlds $2 --steps $3 --settle $4 --variants "fused:1:16" 2>&1 | cut -c1-70; }
for v in product ${VARIANTS:-lit base}; do
  if [ $v = product ]; then unset PS_B200_LIB; else export PS_B200_LIB=$PWD/tools/variants/libps_$v.so; fi
  echo "== $v"; run c2 4096 30 210; run c4 8192 20 150; run c2 1024 30 210; run c3 2048 10 210
done
unset PS_B200_LIB
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
