mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_round2_features.py -m gpu -x -q > gpurun_out/halves_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/halves.rc
( timeout -s KILL 120 python tools/halves_timeline.py deep_sea/11 65536 api fence
  BSB_HOST_SPLIT=0 timeout -s KILL 120 python tools/halves_timeline.py deep_sea/11 65536 api fence ) > gpurun_out/halves_timeline2.txt 2>&1
timeout -s KILL 400 python bench.py --steps 100 --warmup 5 --skip-configs --skip-traffic > gpurun_out/halves_bench.log 2> gpurun_out/halves_bench.err; echo "bench rc=$?" >> gpurun_out/halves.rc
tail -5 gpurun_out/halves_tests.log; cat gpurun_out/halves.rc; cat gpurun_out/halves_timeline2.txt
python - <<'PY'
import json
r = json.loads(open('gpurun_out/halves_bench.log').read().strip().splitlines()[-1])
e = r['e2e']
print('value', r['value'], 'ms', r['ms_per_step'])
print('e2e', e['value'], e['mode']); print('one', e['one_batch_value'], e['one_batch_windows']); print('two', e['two_halves_value'], e['two_halves_windows'])
print('prelaunch', e['prelaunch_value'], 'pipelined', e['pipelined_value'])
PY
