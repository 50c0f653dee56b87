#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched bsuite engine on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU algorithm on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...           # one rank per GPU (weak scaling)

Workload (BASELINE.json configs[1]): deep_sea size=32 (bsuite_id deep_sea/11), 65 536 lanes per GPU, uniform
random actions.  One "step" = one lock-step `step()` call over the whole batch = ONE kernel launch that writes a
fresh dense [B, 32, 32] float32 observation tensor (268 MB) plus reward / discount / step_type.

  value   : env-steps/s with actions already resident in HBM; outputs go to a ring of 4 buffer sets (1.07 GB of
            observations > 126 MB L2, so every step's stores reach HBM); CUDA-event timed, max over ranks.
  e2e     : the same metric through the public Python API with HOST actions (pinned) copied H2D inside the timed
            region and reward / discount / step_type copied D2H every step (observations stay on the device for
            the agent, which is the engine's contract); `host_obs_value` additionally copies the 268 MB of
            observations to pinned host memory every step (PCIe-bound).
  roofline: algorithmic bytes per launch (4 120 B per lane-step) / mean launch duration, against
            MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: the numpy restatement of the reference (oracle/bsuite_oracle.py, kind "port") stepping the same
            workload on all host cores for a bounded sample.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

BSUITE_ID = 'deep_sea/11'          # size = 32, mapping_seed = 42
SIZE = 32
BATCH_PER_GPU = 65536
ALGO_BYTES_PER_LANE_STEP = 4 * SIZE * SIZE + 4 + 4 + 4 + 4 + 4 + 4   # obs + action + reward + discount + step_type + state rd/wr
RING = 4
METRIC = 'env-steps/sec'
FALLBACK_HBM_GBS = 6650.0


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def _reference_worker(job):
  """Steps `lanes` independent oracle environments `warmup + steps` times; returns the timed seconds."""
  lanes, steps, warmup, first_lane = job
  import numpy as np
  from oracle import bsuite_oracle as oracle
  envs = [oracle.OracleEnv('deep_sea', dict(size=SIZE, mapping_seed=42), rng='philox', seed=0, lane=first_lane + i)
          for i in range(lanes)]
  actions = np.random.RandomState(first_lane).randint(2, size=(warmup + steps, lanes))
  for t in range(warmup):
    row = actions[t]
    for i, env in enumerate(envs):
      env.step(int(row[i]))
  start = time.perf_counter()
  for t in range(warmup, warmup + steps):
    row = actions[t]
    for i, env in enumerate(envs):
      env.step(int(row[i]))
  return time.perf_counter() - start


def _calibrate_reference() -> float:
  """Seconds per single-environment step() of the oracle port (one core)."""
  import numpy as np
  from oracle import bsuite_oracle as oracle
  env = oracle.OracleEnv('deep_sea', dict(size=SIZE, mapping_seed=42), rng='philox', seed=0, lane=0)
  actions = np.random.RandomState(0).randint(2, size=3000)
  for a in actions[:500]:
    env.step(int(a))
  start = time.perf_counter()
  for a in actions[500:]:
    env.step(int(a))
  return (time.perf_counter() - start) / 2500


def run_reference_sample(steps: int, warmup: int, budget_s: float = 20.0):
  """The reference's algorithm on every host core, one process per core (mirrors baselines/utils/pool.py:28-54)."""
  import multiprocessing as mp
  cores = os.cpu_count() or 1
  per_step = _calibrate_reference()
  if steps <= 0:   # auto: the whole batch split over the cores, as many steps as fit the time budget
    lanes = max(1, BATCH_PER_GPU // cores)
    steps = max(10, int(budget_s / (lanes * per_step)))
  else:
    lanes = int(budget_s / ((steps + warmup) * per_step))
    lanes = max(1, min(lanes, BATCH_PER_GPU // cores))
  jobs = [(lanes, steps, warmup, w * lanes) for w in range(cores)]
  with mp.get_context('spawn').Pool(cores) as pool:
    seconds = pool.map(_reference_worker, jobs)
  total_steps = cores * lanes * steps
  slowest = max(seconds)
  return dict(value=total_steps / slowest, cores=cores, lanes=cores * lanes, steps=steps, seconds=slowest,
              single_core_steps_per_s=1.0 / per_step)


def reference_main(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return 0
  budget = float(os.environ.get('BSB_BENCH_BUDGET_S', '12.0' if args.steps <= 0 else '20.0'))
  r = run_reference_sample(args.steps, args.warmup, budget_s=budget)
  args.steps = r['steps']
  sample = (f"{r['lanes']} of {BATCH_PER_GPU} lanes x {r['steps']} steps, one process per core "
            f"({r['cores']} cores), numpy restatement of bsuite DeepSea.step")
  line = {
      'metric': METRIC, 'value': r['value'], 'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': 1e3 * r['seconds'] / r['steps'], 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'impl': 'reference',
      'config': {'workload': f'deep_sea size={SIZE} ({BSUITE_ID}) uniform random actions, CPU sample of the '
                             f'{BATCH_PER_GPU}-lane batch', 'sample_lanes': r['lanes']},
      'cpu_baseline': {'value': r['value'], 'unit': 'env-steps/s', 'cores': r['cores'], 'kind': 'port',
                       'sample': sample, 'single_core': r['single_core_steps_per_s']},
      'e2e': {'value': r['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line))
  return 0


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons; `stop(t0, t1)` keeps the samples taken under load."""
  QUERY = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index: int):
    self.index, self.proc, self.lines = index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                    '-lms', '50', '-i', str(self.index)], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._pump, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def wait_first_sample(self, timeout=5.0):
    end = time.time() + timeout
    while self.proc is not None and not self.lines and time.time() < end:
      time.sleep(0.02)

  def stop(self, t0=None, t1=None):
    import datetime
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons, power = [], [], set(), []
    names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
    for line in self.lines:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 8:
        continue
      try:
        stamp = datetime.datetime.strptime(parts[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
        if t0 is not None and not (t0 <= stamp <= t1):
          continue
        sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
      except ValueError:
        continue
      for name, flag in zip(names, parts[4:8]):
        if flag.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples'], 'raw_lines': len(self.lines)}
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2], 'sm_mhz_min': sm[0], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
            'samples_under_load': len(sm), 'power_w_max': max(power)}


# ----------------------------------------------------------------------------- engine arm
def measured_traffic_bytes():
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the headline kernel, from the committed `ncu --set
  full` capture under profiles/ (tools/extract_ncu.py); None when no capture is committed."""
  import csv
  import glob
  paths = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_deep_sea_bulk_ncu_metrics.csv')))
  if not paths:
    return None, None
  scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
  total = 0.0
  with open(paths[-1]) as fh:
    for row in csv.reader(fh):
      if row and row[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and len(row) > 2:
        total += float(row[2].replace(',', '')) * scale.get(row[1], 1.0)
  return (total or None), os.path.relpath(paths[-1], ROOT)


def engine_main(args):
  import torch
  import torch.distributed as dist
  import bsuite_b200
  from bsuite_b200 import _lib

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise RuntimeError('bench.py measures the CUDA engine; no CUDA device is visible (use --impl reference for the CPU arm)')
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)
  if args.gpus != world:
    raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}')

  # cpu_baseline first (rank 0, N = 1 only), in a clean subprocess so worker processes never inherit CUDA state
  cpu_baseline = None
  if rank == 0 and world == 1 and not args.skip_cpu_baseline:
    proc = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '0',
                           '--warmup', '3'], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
    try:
      cpu_baseline = json.loads(proc.stdout.strip().splitlines()[-1])['cpu_baseline']
    except Exception:  # pylint: disable=broad-except
      cpu_baseline = {'value': None, 'unit': 'env-steps/s', 'cores': os.cpu_count(), 'kind': 'port',
                      'sample': 'failed: ' + (proc.stderr or '')[-300:]}

  # For context only: the engine's own explicit host path (the same transition functions compiled for the CPU,
  # one thread) on a 4 096-lane slice of the workload.
  if cpu_baseline is not None and cpu_baseline.get('value'):
    host_env = bsuite_b200.load_from_id(BSUITE_ID, batch=4096, device='cpu', seed=0)
    host_buf = host_env.make_buffers()
    host_act = torch.randint(0, 2, (8, 4096), dtype=torch.int32)
    for t in range(3):
      host_env.step(host_act[t], out=host_buf)
    t0 = time.perf_counter()
    n_host = 0
    while time.perf_counter() - t0 < 1.5:
      host_env.step(host_act[n_host % 8], out=host_buf)
      n_host += 1
    cpu_baseline['engine_host_path_1core'] = 4096 * n_host / (time.perf_counter() - t0)
    host_env.close()

  B, K, W = BATCH_PER_GPU, args.steps, args.warmup
  lib = _lib.load()
  env = bsuite_b200.load_from_id(BSUITE_ID, batch=B, device=device, seed=0, lane_offset=rank * B,
                                 track_episodes=not args.no_track)
  ring = [env.make_buffers() for _ in range(RING)]
  gen = torch.Generator(device=device)
  gen.manual_seed(1234 + rank)
  actions = torch.randint(0, 2, (W + K, B), generator=gen, device=device, dtype=torch.int32)

  def log_point():
    """Device-side reduction of the Logging accumulators + one all-gather of per-rank episode returns."""
    if args.no_track:
      return None
    sums = env.episode_stat_sums()          # one reduction kernel: (steps, episode, total_return, len, return)
    block = sums[[2, 1, 0]]
    if world > 1:
      gathered = torch.empty(world * 3, dtype=block.dtype, device=device)
      dist.all_gather_into_tensor(gathered, block)
      return gathered
    return block

  # ---- value: device-resident actions -------------------------------------
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
    sampler.wait_first_sample()

  def keep_busy(seconds):
    """Untimed steps of the same workload, so the clock samples bracket the timed region under load."""
    end = time.time() + seconds
    t = 0
    while time.time() < end:
      for _ in range(50):
        env.step(actions[t % (W + K)], out=ring[t % RING])
        t += 1
      torch.cuda.synchronize()

  load_t0 = time.time()
  for t in range(W):
    env.step(actions[t], out=ring[t % RING])
  keep_busy(0.5)
  log_point()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  launches0 = lib.bsb_launch_count()
  ev0, ev1, ev2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
  ev0.record()
  for t in range(K):
    env.step(actions[W + t], out=ring[t % RING])
  ev1.record()
  summary = log_point()
  ev2.record()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  launches = lib.bsb_launch_count() - launches0
  keep_busy(0.4)
  clocks = sampler.stop(load_t0 + 0.15, time.time()) if rank == 0 else None
  step_ms = ev0.elapsed_time(ev1)        # K kernel launches back to back
  total_ms = ev0.elapsed_time(ev2)       # + the log point (reduction, all-gather)
  times = torch.tensor([total_ms, step_ms], dtype=torch.float64, device=device)
  if world > 1:
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
  total_ms, step_ms = float(times[0]), float(times[1])
  value = world * B * K / (total_ms * 1e-3)

  # ---- e2e: host actions in, scalars out, every step ---------------------------
  # The public host-buffer call (BatchedEnvironment.step_host -> bsb_step_host): pinned actions H2D, the kernel,
  # reward / discount / step_type D2H, one stream synchronise -- the agent reads the result before acting again.
  Ke = max(10, min(K, 200))
  host_actions = torch.randint(0, 2, (Ke, B), dtype=torch.int32).pin_memory()
  host_small = env.make_host_buffers(with_observation=False)

  def e2e_loop(n, host):
    for t in range(n):
      env.step_host(host_actions[t % Ke], host, out=ring[t % RING])

  def timed_e2e(n, host):
    e2e_loop(3, host)
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_loop(n, host)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
      dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return world * B * n / float(dt[0])

  e2e_value = timed_e2e(Ke, host_small)

  # The same host-memory traffic WITHOUT a host synchronise per step (actions that do not depend on the previous
  # result, as in this random-action workload): env.step() given a pinned host action tensor and outputs whose
  # scalars live in pinned host memory -- the kernel reads / writes them in place; one synchronise at the end.
  mixed = [env.make_mixed_buffers() for _ in range(RING)]
  action_rows = [host_actions[i] for i in range(Ke)]
  for t in range(5):
    env.step(action_rows[t % Ke], out=mixed[t % RING])
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  p0.record()
  for t in range(K):
    env.step(action_rows[t % Ke], out=mixed[t % RING])
  p1.record()
  torch.cuda.synchronize()
  pms = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=device)
  if world > 1:
    dist.all_reduce(pms, op=dist.ReduceOp.MAX)
  e2e_pipelined = world * B * K / (float(pms[0]) * 1e-3)
  host_obs_value = None
  if not args.skip_host_obs:
    host_obs_value = timed_e2e(5, env.make_host_buffers(with_observation=True))

  # ---- the T-fused variant (SURVEY.md 8d asks for both): 16 steps per launch, on-device Philox actions -------
  fused = None
  if not args.skip_fused:
    Tf, reps = 16, 8
    fbuf = [env.make_buffers(Tf) for _ in range(2)]          # 2 x 4.3 GB of observations
    for i in range(2):
      env.rollout(Tf, out=fbuf[i % 2])
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(reps):
      env.rollout(Tf, out=fbuf[i % 2])
    f1.record()
    torch.cuda.synchronize()
    fms = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=device)
    if world > 1:
      dist.all_reduce(fms, op=dist.ReduceOp.MAX)
    per_step_s = float(fms[0]) * 1e-3 / (reps * Tf)
    fused = {'T': Tf, 'value': world * B / per_step_s, 'unit': 'env-steps/s', 'us_per_step': per_step_s * 1e6,
             'achieved_gbs': ALGO_BYTES_PER_LANE_STEP * B / per_step_s / 1e9,
             'note': 'bsb_rollout: 16 steps per launch, lane state in registers, actions sampled on device'}
    del fbuf

  # ---- the same single-step launches replayed from a CUDA graph ("graph-captured") ------------
  graph_replay = None
  if not args.skip_graph and world == 1:      # a per-GPU figure; the multi-rank runs measure scaling, not this
    reps, g_ms, g_err = max(1, K // RING), 0.0, None
    try:     # an optional leg must not take the line down (and holds no collective, so no rank can strand another)
      genv = bsuite_b200.load_from_id(BSUITE_ID, batch=B, device=device, seed=0, lane_offset=rank * B,
                                      track_episodes=not args.no_track)
      graphed = genv.capture(RING)          # RING launches per graph, each writing its own buffer set (> L2 in total)
      graphed.actions.copy_(actions[:RING])
      for _ in range(3):
        graphed.replay()
      torch.cuda.synchronize()
      g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      g0.record()
      for _ in range(reps):
        graphed.replay()
      g1.record()
      torch.cuda.synchronize()
      g_ms = g0.elapsed_time(g1)
      del graphed
      genv.close()
    except Exception as exc:  # pylint: disable=broad-except
      g_err = repr(exc)[:300]
    gms = torch.tensor([g_ms, 0.0 if g_err is None else 1.0], dtype=torch.float64, device=device)
    if world > 1:
      dist.all_reduce(gms, op=dist.ReduceOp.MAX)
    if float(gms[1]) > 0:
      graph_replay = {'value': None, 'error': g_err or 'failed on another rank'}
    else:
      per_step_s = float(gms[0]) * 1e-3 / (reps * RING)
      graph_replay = {'value': world * B / per_step_s, 'unit': 'env-steps/s', 'us_per_step': per_step_s * 1e6,
                      'steps_per_graph': RING, 'replays': reps,
                      'note': 'cudaGraphLaunch of RING captured single-step launches; step counter and chunk scheduler '
                              'live in device memory (graph-safe mode); programmatic edges between the captured '
                              'launches; ranks are not barrier-aligned for this leg'}

  if rank == 0:
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
      peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (measured copy)'
    else:
      peak, peak_src = FALLBACK_HBM_GBS, 'fallback 6.65 TB/s (B200_PROFILING.md)'
    launch_s = (step_ms * 1e-3) / K
    achieved = ALGO_BYTES_PER_LANE_STEP * B / launch_s / 1e9
    traffic, traffic_src = measured_traffic_bytes()
    line = {
        'metric': METRIC, 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': total_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u32 lane state, f64 reward, f32 observation', 'data': 'synthetic',
        'config': {'workload': f'deep_sea size={SIZE} batch={B} per GPU ({BSUITE_ID}), uniform random actions',
                   'bsuite_id': BSUITE_ID, 'batch_per_gpu': B, 'global_batch': world * B,
                   'parallelism': f'lanes sharded over {world} GPU(s), no data-path collective; one all-gather of '
                                  'per-rank episode returns at the log point inside the timed region',
                   'l2_policy': f'outputs cycle through {RING} buffer sets ({RING * B * SIZE * SIZE * 4 / 1e9:.2f} GB '
                                'of observations > 126 MB L2)',
                   'launch': 'value / roofline: one transition_kernel launch per step (T = 1, programmatic dependent launch); the T-fused variant is reported under fused_rollout',
                   'track_episodes': not args.no_track},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': ALGO_BYTES_PER_LANE_STEP * B,
                     'launch_us': launch_s * 1e6, 'kernel': 'transition_kernel<DeepSea, Philox, no-noise, track>: persistent grid, TMA bulk stores of 8 tiles (32 KB)'},
        'cpu_baseline': cpu_baseline,
        'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 4 * B, 'd2h_bytes_per_step': 12 * B,
                'steps': Ke, 'host_obs_value': host_obs_value, 'pipelined_value': e2e_pipelined,
                'host_obs_d2h_bytes_per_step': 4 * B * SIZE * SIZE + 12 * B,
                'note': 'BatchedEnvironment.step_host -> bsb_step_host every step: actions come from pinned host memory and '
                        'reward/discount/step_type land in pinned host memory (read / written in place over PCIe by the '
                        'kernel: zero-copy), then a stream synchronise; observations stay on the device (the API '
                        'contract). host_obs_value also copies the observations to pinned host memory every step. '
                        'pipelined_value: the same per-step host traffic through env.step() with pinned actions and '
                        'pinned scalar outputs, launches queued, one synchronise at the end.'},
        'gpu_launches': int(launches),
        'fused_rollout': fused,
        'graph_replay': graph_replay,
        'clocks': clocks,
        'log_point': None if summary is None else [float(x) for x in summary.cpu()[:3]],
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()
  return 0


def main():
  parser = argparse.ArgumentParser()
  parser.add_argument('--gpus', type=int, default=1)
  parser.add_argument('--steps', type=int, default=400)
  parser.add_argument('--warmup', type=int, default=20)
  parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  parser.add_argument('--skip-cpu-baseline', action='store_true')
  parser.add_argument('--skip-host-obs', action='store_true')
  parser.add_argument('--no-track', action='store_true', help='disable the per-lane Logging accumulators')
  parser.add_argument('--skip-fused', action='store_true', help='skip the T-fused rollout variant')
  parser.add_argument('--skip-graph', action='store_true', help='skip the CUDA-graph replay variant')
  args = parser.parse_args()
  if args.warmup < 3:
    args.warmup = 3
  if args.impl == 'reference':
    return reference_main(args)
  return engine_main(args)


if __name__ == '__main__':
  sys.exit(main())
