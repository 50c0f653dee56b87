#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
BSB_DEEP_SEA_BULK=1 timeout 600 python -m pytest tests -m gpu -x -q -k "deep_sea" > gpurun_out/pytest_dsbulk.log 2>&1; tail -3 gpurun_out/pytest_dsbulk.log
BSB_DEEP_SEA_BULK=1 BSB_DEEP_SEA_PERSISTENT=0 timeout 600 python -m pytest tests -m gpu -x -q -k "deep_sea" > gpurun_out/pytest_dsbulk2.log 2>&1; tail -3 gpurun_out/pytest_dsbulk2.log
timeout 900 python tools/bench_variants.py --out gpurun_out/variants.jsonl > gpurun_out/variants.log 2>&1; echo "rc=$?" >> gpurun_out/variants.log
cat gpurun_out/variants.log
timeout 300 python bench.py --steps 400 --warmup 20 --skip-cpu-baseline > gpurun_out/bench_vec.log 2> gpurun_out/bench.err; tail -n 1 gpurun_out/bench_vec.log | cut -c1-200
BSB_DEEP_SEA_BULK=1 timeout 300 python bench.py --steps 400 --warmup 20 --skip-cpu-baseline > gpurun_out/bench_bulk.log 2>> gpurun_out/bench.err; tail -n 1 gpurun_out/bench_bulk.log | cut -c1-200
BSB_DEEP_SEA_BULK=1 timeout 300 python bench.py --steps 400 --warmup 20 --skip-cpu-baseline --no-track > gpurun_out/bench_bulk_notrack.log 2>> gpurun_out/bench.err; tail -n 1 gpurun_out/bench_bulk_notrack.log | cut -c1-200
tail -3 gpurun_out/bench.err
