#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
BSB_DEEP_SEA_BULK=1 timeout 600 python -m pytest tests -m gpu -x -q -k "deep_sea" > gpurun_out/pytest_dsbulk.log 2>&1; tail -3 gpurun_out/pytest_dsbulk.log
BSB_EMIT_BULK=0 BSB_PDL=0 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_nobulk.log 2>&1; tail -3 gpurun_out/pytest_nobulk.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python tools/bench_variants.py --out gpurun_out/variants.jsonl > gpurun_out/variants.log 2>&1; echo "rc=$?" >> gpurun_out/variants.log
cat gpurun_out/variants.log
timeout 900 python tools/bench_families.py --out gpurun_out/families_bulk.jsonl > gpurun_out/families_bulk.log 2>&1; cat gpurun_out/families_bulk.log
BSB_EMIT_BULK=0 timeout 900 python tools/bench_families.py --out gpurun_out/families_vec.jsonl > gpurun_out/families_vec.log 2>&1; cat gpurun_out/families_vec.log
timeout 300 python bench.py --steps 400 --warmup 20 --skip-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -n 2 gpurun_out/bench.log gpurun_out/bench.err
