#!/bin/bash
# One gpurun call that re-validates everything this repo claims on a B200:
#   parity tests, smoke, bench (+ reference arm), family table, ncu launch list and a full capture of the headline kernel.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh'
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 2000 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err; tail -n 1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 400 --warmup 3 > gpurun_out/bench_reference.log 2>> gpurun_out/bench.err; tail -n 1 gpurun_out/bench_reference.log | cut -c1-300
timeout 900 python tools/bench_families.py --out gpurun_out/families.jsonl > gpurun_out/families.log 2>&1; cat gpurun_out/families.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 40 --warmup 3 --skip-cpu-baseline --skip-host-obs --skip-fused --skip-graph > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 30 -c 2 -o gpurun_out/prof_deep_sea \
  python bench.py --steps 40 --warmup 3 --skip-cpu-baseline --skip-host-obs --skip-fused --skip-graph > gpurun_out/ncu_full.log 2>&1
timeout 300 python tools/bench_sweep.py > gpurun_out/sweep.log 2>&1; tail -n 1 gpurun_out/sweep.log | cut -c1-400
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck.log
tail -3 gpurun_out/bench.err
