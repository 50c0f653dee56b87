#!/bin/bash
# ncu --set full captures of one single-step launch and two fused-rollout launches per family.
# Usage (on the GPU box): bash tools/ncu_families.sh [tag]   -> gpurun_out/ncu_<tag>_<family>.ncu-rep
# Launch order inside tools/bench_families.py --steps 6 --rollout 16: 1 constructor launch, 3 + 6 single steps,
# 2 + 6 rollouts; --launch-skip 9 --launch-count 3 therefore keeps the last single step and the first two rollouts.
tag=${1:-r02}
mkdir -p gpurun_out
for fam in "catch/0" "cartpole/0" "mountain_car/0" "mnist/0" "umbrella_distract/22" "memory_size/16" "bandit/0"; do
  name=$(echo "$fam" | tr '/' '_')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:transition_kernel \
    --launch-skip 9 --launch-count 3 -f -o gpurun_out/ncu_${tag}_${name} \
    python tools/bench_families.py --only "$fam" --steps 6 --rollout 16 > gpurun_out/ncu_${tag}_${name}.log 2>&1
  echo "$fam rc=$?"
done
