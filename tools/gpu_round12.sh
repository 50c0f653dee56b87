#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python tools/bench_families.py --out gpurun_out/families.jsonl > gpurun_out/families.log 2>&1; cat gpurun_out/families.log
timeout 600 python bench.py --steps 2000 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err; tail -n 1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['fused_rollout'], d['cpu_baseline'], d['clocks'])"
tail -2 gpurun_out/bench.err
