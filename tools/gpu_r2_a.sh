#!/bin/bash
# Round 2, first GPU session: parity tests (old + new), smoke, the bench line on the new default workload (C3) with its
# parity sample and rollout figure, sanitizer runs, launch list with DRAM bytes of one settled C3 step, per-round times.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2a_smoke.log | cut -c1-160
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; echo "bench c3 rc=$?"; cut -c1-700 gpurun_out/r2a_bench_c3.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2a_bench_ref_c3.json 2> gpurun_out/r2a_bench_ref_c3.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/r2a_bench_ref_c3.json
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "test_c2_c3_slices" > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "test_c2_c3_slices" > gpurun_out/r2a_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2a_racecheck.log
timeout 600 python tools/sweep.py --workload c3 --worlds 16384 --steps 6 --settle 210 --variants "wide:1:16" --prof > gpurun_out/r2a_wide_rounds_c3.txt 2>&1; echo "wide prof rc=$?"; head -3 gpurun_out/r2a_wide_rounds_c3.txt | cut -c1-250
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-extra --rollout 0 --profile-step"
timeout 900 ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2a_launches_c3_step.csv $CMD > gpurun_out/r2a_ncu_c3_step.log 2>&1; echo "ncu c3 step rc=$?"
ls -la gpurun_out | tail -12
