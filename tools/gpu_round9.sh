#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
# compute-sanitizer on small batches (memcheck + racecheck cover the shared-memory stages / bulk stores)
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_golden_parity.py -m gpu -x -q -k "rollout and (deep_sea_32 or catch_noise or cartpole_noise or umbrella_length or mnist_noise or memory_size or bandit_3 or discounting)" > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.log; tail -4 gpurun_out/sanitizer_$tool.log
done
timeout 600 python bench.py --steps 1000 --warmup 20 --skip-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -n 1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['fused_rollout'])"
tail -3 gpurun_out/bench.err
