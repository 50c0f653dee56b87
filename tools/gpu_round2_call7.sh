set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python tools/e2e_timeline.py > gpurun_out/e2e_timeline.log 2>&1; cat gpurun_out/e2e_timeline.log
BSB_HOST_STAGE_ACTIONS=0 timeout 300 python tools/e2e_timeline.py > gpurun_out/e2e_timeline_inplace_actions.log 2>&1; cat gpurun_out/e2e_timeline_inplace_actions.log
timeout 300 python tools/e2e_timeline.py catch/0 131072 > gpurun_out/e2e_timeline_catch.log 2>&1; cat gpurun_out/e2e_timeline_catch.log
timeout 600 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; cat gpurun_out/e2e_breakdown.log
timeout 600 python bench.py --skip-cpu-baseline --skip-configs > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python tools/bench_families.py --graph 16 --out gpurun_out/families.jsonl > gpurun_out/families.log 2>&1; cut -c1-33,100-130,170-260 gpurun_out/families.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
tail -3 gpurun_out/bench.err
