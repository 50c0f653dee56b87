#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_sweep.py --lanes 4096 --steps 64 --iters 20 > gpurun_out/sweep_n1.log 2>&1; tail -n 1 gpurun_out/sweep_n1.log | cut -c1-400
timeout 300 python bench.py --steps 1000 --warmup 20 --skip-cpu-baseline --skip-host-obs > gpurun_out/bench_n1_same_box.log 2> gpurun_out/bench_same_box.err; tail -n 1 gpurun_out/bench_n1_same_box.log | cut -c1-220
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 20 --skip-host-obs > gpurun_out/bench_n2_same_box.log 2>> gpurun_out/bench_same_box.err; tail -n 1 gpurun_out/bench_n2_same_box.log | cut -c1-220
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 1000 --warmup 20 --skip-cpu-baseline --skip-host-obs > gpurun_out/bench_n1_gpu1.log 2>> gpurun_out/bench_same_box.err; tail -n 1 gpurun_out/bench_n1_gpu1.log | cut -c1-220
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_sweep.py --gpus 2 --lanes 4096 --steps 64 --iters 20 > gpurun_out/sweep_n2.log 2>&1; tail -n 1 gpurun_out/sweep_n2.log | cut -c1-300
tail -5 gpurun_out/bench_same_box.err
