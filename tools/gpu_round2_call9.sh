mkdir -p gpurun_out
for bt in 32 64 128; do
  for fam in catch/0 cartpole/0 mountain_car/0 bandit/0 memory_len/5 umbrella_length/10; do
    BSB_BLOCK_THREADS=$bt timeout 120 python tools/bench_families.py --only "$fam" --steps 100 2>&1 | cut -c1-150 | sed "s/^/threads=$bt  /"
  done
done > gpurun_out/ab_block_threads.log 2>&1
cat gpurun_out/ab_block_threads.log
