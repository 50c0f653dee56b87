"""CPU arm of bench.py: the reference's own `step()` loop on the host cores, one process per core.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under bsuite_b200/ imports this module; bench.py runs it for
`--impl reference` and for its `cpu_baseline` leg, never on the product path.

What is timed (north_star: "the reference's own NumPy step() loop timed on the same box's host cores"):

    for action in actions: env.step(action)          # README.md:177-184; environments/deep_sea.py:116-144

on `lanes_per_worker` independent `bsuite.environments.deep_sea.DeepSea(size=32, mapping_seed=42)` objects per
process, one process per usable core (the reference's own parallelism is a process pool over environments:
baselines/utils/pool.py:28-54).  kind = "reference" when the unmodified reference is importable from
`oracle/_ref` (installed by oracle/install_ref.py; travels to the GPU box like the built .so), else kind = "port":
the numpy restatement `oracle/bsuite_oracle.py` (pinned against the reference by tests/test_oracle_pinned.py).

Why it is written this way ("the CPU denominator is unstable by 10x"):
  * BLAS/OpenMP thread pools are pinned to 1 thread per process BEFORE the workers start;
  * the worker count is the USABLE core count (affinity mask and cgroup quota, not just os.cpu_count());
  * every worker builds its environments and warms up first, then all workers enter each timed round through a
    barrier, so a round measures the fully loaded machine and nothing else;
  * a round lasts a fixed wall time (>= 2 s) and at least `min_passes` passes; the figure is
    sum(lane-steps) / max(elapsed) and the MEDIAN of `rounds` rounds is reported.
"""

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, '_ref')
SHIMS = os.path.join(HERE, 'shims')
SIZE = 32

_THREAD_VARS = ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS',
                'VECLIB_MAXIMUM_THREADS')


def reference_installed() -> bool:
  return os.path.isfile(os.path.join(REF_DIR, 'bsuite', 'environments', 'deep_sea.py'))


def usable_cores():
  """(cores usable by this process, how that was derived)."""
  count = os.cpu_count() or 1
  notes = {'os_cpu_count': count}
  try:
    affinity = len(os.sched_getaffinity(0))
    notes['affinity'] = affinity
    count = min(count, affinity)
  except (AttributeError, OSError):
    pass
  try:
    with open('/sys/fs/cgroup/cpu.max') as fh:
      quota, period = fh.read().split()
    notes['cgroup_cpu_max'] = f'{quota} {period}'
    if quota != 'max':
      count = max(1, min(count, int(float(quota) / float(period))))
  except (OSError, ValueError):
    pass
  return count, notes


def _make_envs(kind: str, lanes: int, first_lane: int):
  if kind == 'reference':
    for path in (SHIMS, REF_DIR):
      if path not in sys.path:
        sys.path.insert(0, path)
    from bsuite.environments import deep_sea  # the UNMODIFIED reference (pip-installed into oracle/_ref)
    return [deep_sea.DeepSea(size=SIZE, seed=first_lane + i, mapping_seed=42) for i in range(lanes)]
  if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
  from oracle import bsuite_oracle as oracle
  return [oracle.OracleEnv('deep_sea', dict(size=SIZE, mapping_seed=42), rng='philox', seed=0, lane=first_lane + i)
          for i in range(lanes)]


def _worker(index, kind, lanes, rounds, seconds, min_passes, barrier, results):
  import numpy as np
  envs = _make_envs(kind, lanes, index * lanes)
  steps = [env.step for env in envs]
  table = np.random.RandomState(index).randint(2, size=(1024, lanes)).tolist()      # Python ints, as an agent passes
  for row in table[:3]:                                                             # warm-up passes
    for step, action in zip(steps, row):
      step(action)
  for r in range(rounds):
    barrier.wait()
    passes, start = 0, time.perf_counter()
    while True:
      row = table[passes & 1023]
      for step, action in zip(steps, row):
        step(action)
      passes += 1
      if passes >= min_passes and (passes & 7) == 0 and time.perf_counter() - start >= seconds:
        break
    results.put((index, r, passes, time.perf_counter() - start))


def run(kind=None, rounds: int = 3, seconds: float = 2.0, min_passes: int = 8, lanes_per_worker: int = 8,
        workers=None):
  """Returns a dict: value (env-steps/s, median round), cores, kind, per-round values, timing and host notes."""
  import multiprocessing as mp
  if kind is None:
    kind = 'reference' if reference_installed() else 'port'
  cores, notes = usable_cores()
  if workers:
    cores = int(workers)
  saved = {k: os.environ.get(k) for k in _THREAD_VARS}
  for k in _THREAD_VARS:          # inherited by the spawned workers: one thread per process, whatever launched us
    os.environ[k] = '1'
  try:
    ctx = mp.get_context('spawn')
    barrier, results = ctx.Barrier(cores), ctx.Queue()
    min_passes = (int(min_passes) + 7) // 8 * 8
    procs = [ctx.Process(target=_worker, args=(w, kind, lanes_per_worker, rounds, seconds, min_passes, barrier, results),
                         daemon=True) for w in range(cores)]
    t_start = time.perf_counter()
    for p in procs:
      p.start()
    rows = [results.get(timeout=600) for _ in range(cores * rounds)]
    for p in procs:
      p.join(timeout=30)
    wall = time.perf_counter() - t_start
  finally:
    for k, v in saved.items():
      if v is None:
        os.environ.pop(k, None)
      else:
        os.environ[k] = v
  per_round = []
  for r in range(rounds):
    mine = [row for row in rows if row[1] == r]
    lane_steps = sum(passes for _, _, passes, _ in mine) * lanes_per_worker
    slowest = max(elapsed for _, _, _, elapsed in mine)
    per_round.append(dict(value=lane_steps / slowest, seconds=slowest,
                          passes_min=min(p for _, _, p, _ in mine), passes_max=max(p for _, _, p, _ in mine)))
  ordered = sorted(per_round, key=lambda d: d['value'])
  median = ordered[len(ordered) // 2]
  return dict(value=median['value'], kind=kind, cores=cores, lanes=cores * lanes_per_worker,
              lanes_per_worker=lanes_per_worker, rounds=[d['value'] for d in per_round],
              seconds=median['seconds'], passes=median['passes_min'], wall_seconds=wall, host=notes,
              per_core=median['value'] / cores)


if __name__ == '__main__':
  import json
  print(json.dumps(run(rounds=3, seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 2.0)))
