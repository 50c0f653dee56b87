# Final GPU call of round 2 (~8.8 GPU-minutes left, ~5.5 of them run time): the driver's line with the HostParts legs,
# the new tests, the whole GPU suite, two A/B variants, timelines of 3 / 4 parts, an ncu launch list of the split step.
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))" >> gpurun_out/last2.rc; }
: > gpurun_out/last2.rc
timeout -s KILL 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last2_bench_k20.json 2> gpurun_out/last2_bench_k20.err; stamp bench $?
timeout -s KILL 200 python -m pytest tests -m gpu -x -q > gpurun_out/last2_tests.log 2>&1; stamp tests $?
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last2_smoke.log 2>&1; stamp smoke $?
( for parts in 3 4; do timeout -s KILL 60 python tools/halves_timeline.py deep_sea/11 65536 api fence $parts; done ) > gpurun_out/last2_parts_timeline.txt 2>&1; stamp timeline $?
QUICK="--steps 100 --warmup 5 --skip-configs --skip-traffic --skip-fused --skip-graph --skip-host-obs --skip-cpu-baseline"
BSB_HOST_STAGE_ACTIONS=0 timeout -s KILL 90 python bench.py $QUICK > gpurun_out/last2_bench_nostage.json 2> gpurun_out/last2_bench_nostage.err; stamp nostage $?
BSB_HOST_SPLIT=0 timeout -s KILL 90 python bench.py $QUICK > gpurun_out/last2_bench_nosplit.json 2> gpurun_out/last2_bench_nosplit.err; stamp nosplit $?
timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:transition_kernel --launch-skip 12 --launch-count 16 \
  --csv --log-file gpurun_out/last2_split_launches.csv python tools/parts_probe.py 2 > gpurun_out/last2_ncu.log 2>&1; stamp ncu $?
cat gpurun_out/last2.rc; tail -3 gpurun_out/last2_tests.log; tail -1 gpurun_out/last2_smoke.log
python - <<'PY'
import json
for f in ('gpurun_out/last2_bench_k20.json', 'gpurun_out/last2_bench_nostage.json', 'gpurun_out/last2_bench_nosplit.json'):
  try:
    r = json.loads(open(f).read().strip().splitlines()[-1])
    e = r['e2e']
    print(f, 'value', r['value'], 'ms', r['ms_per_step'], 'frac', r['roofline']['frac'])
    print('  e2e', e['value'], e.get('mode')); print('  one', e.get('one_batch_value'), 'parts', e.get('parts_values'), e.get('parts_errors'))
  except Exception as ex:
    print(f, 'unreadable', ex)
PY
cat gpurun_out/last2_parts_timeline.txt; tail -25 gpurun_out/last2_split_launches.csv
