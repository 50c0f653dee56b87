mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_round2_features.py -m gpu -x -q > gpurun_out/halves_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/halves.rc
timeout 500 python bench.py --steps 100 --warmup 5 --skip-configs --skip-traffic > gpurun_out/halves_bench.log 2> gpurun_out/halves_bench.err; echo "bench rc=$?" >> gpurun_out/halves.rc
tail -5 gpurun_out/halves_tests.log; cat gpurun_out/halves.rc
python - <<'PY'
import json
r = json.loads(open('gpurun_out/halves_bench.log').read().strip().splitlines()[-1])
e = r['e2e']
print('value', r['value'], 'ms', r['ms_per_step'])
print('e2e', e['value'], e['mode']); print('one', e['one_batch_value'], e['one_batch_windows']); print('two', e['two_halves_value'], e['two_halves_windows'])
print('prelaunch', e['prelaunch_value'], 'pipelined', e['pipelined_value'])
PY
