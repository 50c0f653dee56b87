set -x
mkdir -p gpurun_out
timeout 300 python tools/sweep_breakdown.py 512 > gpurun_out/sweep_breakdown_512.log 2>&1; cat gpurun_out/sweep_breakdown_512.log
timeout 300 python tools/sweep_breakdown.py 4096 > gpurun_out/sweep_breakdown_4096.log 2>&1; cat gpurun_out/sweep_breakdown_4096.log
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-configs > gpurun_out/bench_k20.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
