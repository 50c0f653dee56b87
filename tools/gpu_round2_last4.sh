# Closing validation of the final tree: the whole GPU suite, smoke(), the driver's bench line; then (time permitting)
# the whole sanitizer workload under memcheck.
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "$1 rc=$2 t=$(( $(date +%s) - t0 ))" >> gpurun_out/last4.rc; }
: > gpurun_out/last4.rc
timeout -s KILL 120 python -m pytest tests -m gpu -x -q > gpurun_out/last4_tests.log 2>&1; stamp tests $?
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last4_smoke.log 2>&1; stamp smoke $?
timeout -s KILL 120 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last4_bench_k20.json 2> gpurun_out/last4_bench_k20.err; stamp bench $?
timeout -s KILL 100 compute-sanitizer --tool memcheck python tools/sanitize_check.py > gpurun_out/last4_memcheck_full.log 2>&1; stamp memcheck_full $?
cat gpurun_out/last4.rc; tail -2 gpurun_out/last4_tests.log; tail -1 gpurun_out/last4_smoke.log; tail -3 gpurun_out/last4_memcheck_full.log
python - <<'PY'
import json
r = json.loads(open('gpurun_out/last4_bench_k20.json').read().strip().splitlines()[-1])
e = r['e2e']
print('value', r['value'], 'ms', r['ms_per_step'], 'frac', r['roofline']['frac'], 'e2e', e['value'], e['mode'], e['parts_values'])
PY
