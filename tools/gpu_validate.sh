#!/bin/bash
# One gpurun call that re-validates everything this repo claims on ONE B200:
#   parity tests, smoke, bench (+ reference arm), host-step breakdown and timeline, family table, ncu launch list and a
#   full capture of the headline kernel, compute-sanitizer runs.
# Usage: gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh'
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_k20.log 2>> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>> gpurun_out/bench.err
timeout 600 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; cat gpurun_out/e2e_breakdown.log
timeout 300 python tools/e2e_timeline.py > gpurun_out/e2e_timeline.log 2>&1; cat gpurun_out/e2e_timeline.log
BSB_HOST_STAGE_ACTIONS=1 timeout 300 python tools/e2e_timeline.py > gpurun_out/e2e_timeline_dma_actions.log 2>&1; cat gpurun_out/e2e_timeline_dma_actions.log
timeout 300 python tools/e2e_timeline.py catch/0 131072 > gpurun_out/e2e_timeline_catch.log 2>&1; cat gpurun_out/e2e_timeline_catch.log
timeout 900 python tools/bench_families.py --graph 16 --out gpurun_out/families.jsonl > gpurun_out/families.log 2>&1; cut -c1-150 gpurun_out/families.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 40 --warmup 3 --skip-cpu-baseline --skip-host-obs --skip-fused --skip-graph --skip-configs --skip-traffic > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 30 -c 2 -f -o gpurun_out/prof_deep_sea \
  python bench.py --steps 40 --warmup 3 --skip-cpu-baseline --skip-host-obs --skip-fused --skip-graph --skip-configs --skip-traffic > gpurun_out/ncu_full.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
tail -5 gpurun_out/bench.err
