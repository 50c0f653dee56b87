#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_variants.py --out gpurun_out/variants.jsonl > gpurun_out/variants.log 2>&1; echo "rc=$?" >> gpurun_out/variants.log
cat gpurun_out/variants.log
timeout 300 python bench.py --steps 400 --warmup 20 --skip-cpu-baseline > gpurun_out/bench_vec128.log 2> gpurun_out/bench_vec128.err
BSB_BLOCK_THREADS=64 timeout 300 python bench.py --steps 400 --warmup 20 --skip-cpu-baseline > gpurun_out/bench_vec64.log 2> gpurun_out/bench_vec64.err
BSB_DEEP_SEA_EMIT=tma timeout 300 python bench.py --steps 400 --warmup 20 --skip-cpu-baseline > gpurun_out/bench_tma.log 2> gpurun_out/bench_tma.err
tail -n 2 gpurun_out/bench_*.log gpurun_out/bench_*.err
BSB_DEEP_SEA_EMIT=tma timeout 600 python -m pytest tests -m gpu -x -q -k "deep_sea" > gpurun_out/pytest_tma.log 2>&1; tail -3 gpurun_out/pytest_tma.log
BSB_DEEP_SEA_EMIT=tma timeout 600 ncu --set full --clock-control none --import-source on -k regex:deep_sea_bulk -s 6 -c 1 -o gpurun_out/prof_deep_sea_tma \
  python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-host-obs > gpurun_out/ncu_tma.log 2>&1
