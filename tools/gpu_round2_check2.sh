mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py > gpurun_out/multi_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/multi_check_n2.log; grep -v "^  File\|^    \|^$" gpurun_out/multi_check_n2.log | head -40 | cut -c1-600
