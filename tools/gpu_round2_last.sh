# Last GPU call of round 2 (~14 GPU-minutes left): validate HEAD, then the driver's line, then the split A/B.
mkdir -p gpurun_out
t0=$(date +%s)
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/last_tests.log 2>&1; echo "tests rc=$? t=$(( $(date +%s) - t0 ))" > gpurun_out/last.rc
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - t0 ))" >> gpurun_out/last.rc
timeout -s KILL 420 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last_bench_k20.json 2> gpurun_out/last_bench_k20.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))" >> gpurun_out/last.rc
BSB_HOST_SPLIT=0 timeout -s KILL 200 python bench.py --steps 100 --warmup 5 --skip-configs --skip-traffic > gpurun_out/last_bench_nosplit.json 2> gpurun_out/last_bench_nosplit.err; echo "nosplit rc=$? t=$(( $(date +%s) - t0 ))" >> gpurun_out/last.rc
( timeout -s KILL 100 python tools/halves_timeline.py deep_sea/11 65536 api fence
  BSB_HOST_SPLIT=0 timeout -s KILL 100 python tools/halves_timeline.py deep_sea/11 65536 api fence ) > gpurun_out/last_halves_timeline.txt 2>&1; echo "timeline rc=$? t=$(( $(date +%s) - t0 ))" >> gpurun_out/last.rc
tail -3 gpurun_out/last_tests.log; cat gpurun_out/last.rc; tail -2 gpurun_out/last_smoke.log
python - <<'PY'
import json
for f in ('gpurun_out/last_bench_k20.json', 'gpurun_out/last_bench_nosplit.json'):
  try:
    r = json.loads(open(f).read().strip().splitlines()[-1])
    e = r['e2e']
    print(f, 'value', r['value'], 'ms', r['ms_per_step'], 'frac', r['roofline']['frac'])
    print('  e2e', e['value'], e.get('mode')); print('  one', e.get('one_batch_value')); print('  two', e.get('two_halves_value'), e.get('two_halves_windows'))
  except Exception as ex:
    print(f, 'unreadable', ex)
PY
cat gpurun_out/last_halves_timeline.txt
