mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/final_ref_n2.log 2> gpurun_out/final_n2.err; echo "rc=$?" >> gpurun_out/final_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final_n2.log 2>> gpurun_out/final_n2.err; echo "rc=$?" >> gpurun_out/final_n2.err
timeout 500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref_n1.log 2> gpurun_out/final_n1.err; echo "rc=$?" >> gpurun_out/final_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_n1.log 2>> gpurun_out/final_n1.err; echo "rc=$?" >> gpurun_out/final_n1.err
grep rc= gpurun_out/final_n2.err gpurun_out/final_n1.err
