# Usage: bash tools/gpu_round2_multi.sh N      (on a box with N GPUs)
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tools/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/multi_check_n$N.log; tail -3 gpurun_out/multi_check_n$N.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/scale_n1.log 2> gpurun_out/scale_n1.err; echo "rc=$?" >> gpurun_out/scale_n1.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.log 2> gpurun_out/scale_n$N.err; echo "rc=$?" >> gpurun_out/scale_n$N.err
timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_ref_n$N.log 2>> gpurun_out/scale_n$N.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/scale_ref_n1.log 2>> gpurun_out/scale_n1.err
cat /sys/fs/cgroup/cpu.max > gpurun_out/host_n$N.txt; nproc >> gpurun_out/host_n$N.txt
tail -3 gpurun_out/scale_n$N.err gpurun_out/scale_n1.err
