#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 1000 --warmup 20 --skip-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -n 1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['fused_rollout']['value'])"
timeout 600 python bench.py --steps 1000 --warmup 20 --skip-cpu-baseline --no-track --skip-fused > gpurun_out/bench_nt.log 2>> gpurun_out/bench.err; tail -n 1 gpurun_out/bench_nt.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
tail -3 gpurun_out/bench.err
