set -x
mkdir -p gpurun_out
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu-baseline --skip-fused --skip-host-obs --skip-graph > gpurun_out/scale8b_n1.log 2> gpurun_out/scale8b_n1.err; echo "rc=$?" >> gpurun_out/scale8b_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --skip-fused --skip-host-obs > gpurun_out/scale8b_n8.log 2> gpurun_out/scale8b_n8.err; echo "rc=$?" >> gpurun_out/scale8b_n8.err
tail -3 gpurun_out/scale8b_n8.err gpurun_out/scale8b_n1.err
