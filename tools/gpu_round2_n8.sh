set -x
mkdir -p gpurun_out
TR() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
cat /sys/fs/cgroup/cpu.max > gpurun_out/host_n8.txt; nproc >> gpurun_out/host_n8.txt; nvidia-smi -L >> gpurun_out/host_n8.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py > gpurun_out/multi_check_n8.log 2>&1; echo "rc=$?" >> gpurun_out/multi_check_n8.log; grep -v "^  File\|^    \|^$" gpurun_out/multi_check_n8.log | tail -5 | cut -c1-400
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/scale8_n1.log 2> gpurun_out/scale8_n1.err; echo "rc=$?" >> gpurun_out/scale8_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/scale8_n8.log 2> gpurun_out/scale8_n8.err; echo "rc=$?" >> gpurun_out/scale8_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 --skip-configs --skip-fused --skip-host-obs > gpurun_out/scale8_n2.log 2> gpurun_out/scale8_n2.err; echo "rc=$?" >> gpurun_out/scale8_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 20 --warmup 5 --skip-configs --skip-fused --skip-host-obs > gpurun_out/scale8_n4.log 2> gpurun_out/scale8_n4.err; echo "rc=$?" >> gpurun_out/scale8_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/scale8_ref_n8.log 2>> gpurun_out/scale8_n8.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/scale8_ref_n1.log 2>> gpurun_out/scale8_n1.err
tail -3 gpurun_out/scale8_n8.err gpurun_out/scale8_n1.err gpurun_out/scale8_n2.err gpurun_out/scale8_n4.err
