# Last GPU seconds of round 2: sanitizers over the split host steps, then an A/B of co-resident observation-only
# launches (BSB_SPLIT_GROUP / BSB_SPLIT_CTAS_PER_SM; defaults = the validated behaviour), then the HostParts tests
# under the best variant.
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))" >> gpurun_out/last3.rc; }
: > gpurun_out/last3.rc
timeout -s KILL 50 compute-sanitizer --tool memcheck python tools/sanitize_check.py --only-split > gpurun_out/last3_memcheck.log 2>&1; stamp memcheck $?
timeout -s KILL 50 compute-sanitizer --tool racecheck python tools/sanitize_check.py --only-split > gpurun_out/last3_racecheck.log 2>&1; stamp racecheck $?
QUICK="--steps 100 --warmup 5 --skip-configs --skip-traffic --skip-fused --skip-graph --skip-host-obs --skip-cpu-baseline"
run_variant() {   # name group ctas
  BSB_SPLIT_GROUP=$2 BSB_SPLIT_CTAS_PER_SM=$3 timeout -s KILL 25 python bench.py $QUICK > gpurun_out/last3_ab_$1.json 2> gpurun_out/last3_ab_$1.err; stamp "ab_$1(group=$2,ctas=$3)" $?
}
run_variant base 0 0
run_variant g4c3 4 3
run_variant g8c2 0 2
run_variant g4c4 4 4
run_variant g2c6 2 6
run_variant g4 4 0
run_variant g8c1 0 1
python - <<'PY' > gpurun_out/last3_ab_summary.txt
import json
best, best_v = None, 0.0
for name, env in (('base', '0 0'), ('g4c3', '4 3'), ('g8c2', '0 2'), ('g4c4', '4 4'), ('g2c6', '2 6'), ('g4', '4 0'), ('g8c1', '0 1')):
  try:
    r = json.loads(open(f'gpurun_out/last3_ab_{name}.json').read().strip().splitlines()[-1])
    e = r['e2e']
    pv = {k: round(v / 1e9, 4) for k, v in (e.get('parts_values') or {}).items()}
    print(f'{name:6s} group/ctas={env}  value {r["value"] / 1e9:.4f}e9  one_batch {e["one_batch_value"] / 1e9:.4f}e9  parts {pv}  errors {e.get("parts_errors")}')
    top = max((e.get('parts_values') or {'0': 0.0}).values())
    if top > best_v:
      best, best_v = (name, env), top
  except Exception as ex:
    print(name, 'unreadable', repr(ex)[:100])
print('best', best, best_v)
open('gpurun_out/last3_best.env', 'w').write(best[1] if best else '0 0')
PY
cat gpurun_out/last3_ab_summary.txt
read G C < gpurun_out/last3_best.env
BSB_SPLIT_GROUP=$G BSB_SPLIT_CTAS_PER_SM=$C timeout -s KILL 60 python -m pytest tests/test_round2_features.py -m gpu -x -q > gpurun_out/last3_tests_best.log 2>&1; stamp "tests_best(group=$G,ctas=$C)" $?
cat gpurun_out/last3.rc; tail -2 gpurun_out/last3_tests_best.log; tail -4 gpurun_out/last3_memcheck.log; tail -4 gpurun_out/last3_racecheck.log
