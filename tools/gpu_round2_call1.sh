set -x
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/host.txt 2>&1; python -c "import os;print(len(os.sched_getaffinity(0)))" >> gpurun_out/host.txt; uptime >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.log
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/bench_k20.log 2>> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_a.log 2>> gpurun_out/bench.err
timeout 200 env OMP_NUM_THREADS=1 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_b.log 2>> gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 400 --warmup 3 > gpurun_out/bench_ref_c.log 2>> gpurun_out/bench.err
bash tools/ncu_families.sh r02a
tail -5 gpurun_out/bench.err
