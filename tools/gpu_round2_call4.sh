set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; cat gpurun_out/e2e_breakdown.log
timeout 300 python tools/e2e_breakdown.py catch/0 131072 > gpurun_out/e2e_breakdown_catch.log 2>&1; cat gpurun_out/e2e_breakdown_catch.log
timeout 600 python bench.py --skip-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-configs > gpurun_out/bench_k20.log 2>> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err
timeout 600 python tools/bench_families.py --graph 16 --out gpurun_out/families.jsonl > gpurun_out/families.log 2>&1; cat gpurun_out/families.log
for v in "BSB_IMAGE_GROUP=4" "BSB_IMAGE_STAGES=2" "BSB_IMAGE_GROUP=4 BSB_IMAGE_STAGES=2"; do env $v timeout 300 python tools/bench_families.py --only mnist > gpurun_out/families_mnist_$(echo $v | tr ' =' '__').log 2>&1; echo "$v: $(cat gpurun_out/families_mnist_$(echo $v | tr ' =' '__').log)"; done
for fam in mnist/0; do name=$(echo $fam | tr '/' '_'); timeout 300 ncu --set full --clock-control none --import-source on -k regex:transition_kernel --launch-skip 8 --launch-count 4 -f -o gpurun_out/ncu_r02c_${name} python tools/bench_families.py --only "$fam" --steps 6 --rollout 16 > gpurun_out/ncu_r02c_${name}.log 2>&1; done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_check.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck.log
tail -5 gpurun_out/bench.err
