#!/bin/bash
# Small-batch chunk policy check (one gpurun call): the GPU tests under the automatic policy and with 8- / 32-lane
# chunks forced for every family, 4 096-lane per-family rates for both, config #5, sanitizer with 8-lane chunks.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/chunk_auto_pytest.log 2>&1; tail -n 2 gpurun_out/chunk_auto_pytest.log
if [ "$1" == "full" ]; then
BSB_CHUNK_LANES=8 timeout 300 python -m pytest tests -m gpu -q > gpurun_out/chunk8_pytest.log 2>&1; tail -n 2 gpurun_out/chunk8_pytest.log
BSB_CHUNK_LANES=32 timeout 300 python -m pytest tests -m gpu -q > gpurun_out/chunk32_pytest.log 2>&1; tail -n 2 gpurun_out/chunk32_pytest.log
BSB_CHUNK_LANES=32 timeout 200 python tools/bench_families.py --batch 4096 --rollout 64 --steps 30 > gpurun_out/families4096_chunk32.log 2>&1
BSB_CHUNK_LANES=8 timeout 200 python tools/bench_families.py --batch 4096 --rollout 64 --steps 30 > gpurun_out/families4096_chunk8.log 2>&1
BSB_CHUNK_LANES=8 timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/sanitizer_chunk8_memcheck.log 2>&1; tail -n 2 gpurun_out/sanitizer_chunk8_memcheck.log
BSB_CHUNK_LANES=8 timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/sanitizer_chunk8_racecheck.log 2>&1; tail -n 2 gpurun_out/sanitizer_chunk8_racecheck.log
fi
timeout 200 python tools/bench_families.py --batch 4096 --rollout 64 --steps 30 > gpurun_out/families4096_auto.log 2>&1; grep -i "deep_sea\|mnist" gpurun_out/families4096_auto.log
timeout 200 python tools/bench_sweep.py > gpurun_out/sweep_auto.log 2>&1; tail -n 1 gpurun_out/sweep_auto.log | cut -c1-260
timeout 300 python bench.py --steps 2000 --warmup 20 --skip-cpu-baseline --skip-host-obs > gpurun_out/bench_chunk.log 2> gpurun_out/bench_chunk.err; tail -n 1 gpurun_out/bench_chunk.log | cut -c1-200
