#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "step_host or full_size or features" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 1000 --warmup 20 --skip-cpu-baseline --skip-fused > gpurun_out/bench_zc.log 2> gpurun_out/bench.err; tail -n 1 gpurun_out/bench_zc.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('zero-copy', d['value'], d['e2e'])"
BSB_ZERO_COPY=0 timeout 600 python bench.py --steps 1000 --warmup 20 --skip-cpu-baseline --skip-fused > gpurun_out/bench_copy.log 2>> gpurun_out/bench.err; tail -n 1 gpurun_out/bench_copy.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('staged   ', d['value'], d['e2e'])"
tail -3 gpurun_out/bench.err
