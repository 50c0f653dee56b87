#!/bin/bash
# One gpurun call: parity tests, smoke, bench, family table, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -2 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 900 python tools/bench_families.py --out gpurun_out/families.jsonl > gpurun_out/families.log 2>&1
tail -20 gpurun_out/families.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 20 --warmup 3 --skip-cpu-baseline --skip-host-obs > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transition_kernel -s 6 -c 2 -o gpurun_out/prof_deep_sea \
  python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-host-obs > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
