#!/bin/bash
# One GPU session: tests, smoke, benches, ncu launch lists.  (The --set full capture of the step kernel is a separate
# call: tools/gpu_ncu_full.sh.)  Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo "bench ref rc=$?"
timeout 1200 python bench.py --steps 5 --warmup 3 --workload c3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu --settle 20 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --settle 100"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches c2 rc=$?"
CMD3="python bench.py --steps 1 --warmup 3 --no-cpu --settle 100 --workload c3"
$CMD3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 12900 -c 400 --csv --log-file gpurun_out/launches_c3.csv $CMD3 > gpurun_out/ncu_launches_c3.log 2>&1
echo "ncu launches c3 rc=$?"
timeout 300 python tools/stress_fast_vs_serial.py --workload c3 --worlds 96 --steps 200 --path wide 2>&1 | tail -1
timeout 300 python tools/stress_fast_vs_serial.py --workload c2 --worlds 128 --steps 300 --path fused 2>&1 | tail -1
ls -la gpurun_out | tail -12
