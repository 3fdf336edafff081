#!/bin/bash
# One GPU session: tests, smoke, benches, ncu launch lists and the --set full capture of the C2 step kernel.
# Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-120
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo "bench ref rc=$?"
timeout 1200 python bench.py --steps 5 --warmup 3 --workload c3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --workload c4 --no-cpu --settle 20 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
PS_B200_HOST_CHUNKS=16 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2_chunks16.json 2>/dev/null; echo "bench c2 chunks16 rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --settle 100"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches c2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/prof_c2_final -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full c2 rc=$?"
timeout 300 python tools/stress_fast_vs_serial.py --workload c3 --worlds 96 --steps 200 --path wide 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/stress_fast_vs_serial.py --workload c2 --worlds 128 --steps 300 --path fused 2>&1 | tail -1 | cut -c1-200
timeout 200 python tools/bench_c1.py 2>&1 | tail -1 | cut -c1-300
ls -la gpurun_out | tail -5
