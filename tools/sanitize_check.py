#!/usr/bin/env python
"""A workload for compute-sanitizer that exercises the PERSISTENT deep_sea path (dynamic chunk counter, grouped
TMA bulk stores with L2 hints, PDL) plus the catch / row bulk emitters at a size the sanitizer finishes quickly.

    compute-sanitizer --tool memcheck  python tools/sanitize_check.py
    compute-sanitizer --tool racecheck python tools/sanitize_check.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bsuite_b200


def split_steps():
  """Split host steps (BSB_HOST_NO_WAIT: transitions + copiers, then the observation-only launch that waits for the
  phase-1 flag) on 2 / 3 handles driven round-robin, against one ordinary environment."""
  from bsuite_b200 import rollouts
  for bsuite_id, batch, parts in (('deep_sea/11', 12000 + 5, 3), ('deep_sea_stochastic/3', 9000, 2)):
    group = rollouts.HostParts(bsuite_id, batch, device='cuda', seed=3, track_episodes=True, parts=parts)
    twin = bsuite_b200.load_from_id(bsuite_id, batch=batch, device='cuda', seed=3, track_episodes=True)
    bounds = [0]
    for size in group.sizes:
      bounds.append(bounds[-1] + size)
    T = 5
    acts = torch.as_tensor(twin.random_actions(T, action_seed=1, first_step=0))
    rows = [acts[:, bounds[p]:bounds[p + 1]].contiguous().pin_memory() for p in range(parts)]
    group.reset(); twin.reset()
    want = [twin.step(acts[t].cuda(), out=twin.make_buffers()) for t in range(T)]
    torch.cuda.synchronize()

    def check(p, t, ts, obs):
      lanes = slice(bounds[p], bounds[p + 1])
      torch.cuda.synchronize()
      assert torch.equal(obs, want[t].observation[lanes]) and torch.equal(ts.reward, want[t].reward.cpu()[lanes]) \
          and torch.equal(ts.step_type, want[t].step_type.cpu()[lanes]), (p, t)

    for p in range(parts):
      group.submit(p, rows[p][0])
    for t in range(1, T):
      for p in range(parts):
        check(p, t - 1, *group.collect(p)); group.submit(p, rows[p][t])
    for p in range(parts):
      check(p, T - 1, *group.collect(p))
    print(bsuite_id, batch, f'{parts} parts, split host steps == ordinary steps: True', flush=True)
    group.close(); twin.close()


if '--only-split' in sys.argv:
  split_steps()
  print('sanitize workload (split steps) finished')
  sys.exit(0)

for bsuite_id, batch in (('deep_sea/11', 20000), ('deep_sea_stochastic/3', 40001), ('catch_noise/0', 5000), ('cartpole/0', 3000),
                         ('umbrella_length/3', 2000), ('umbrella_distract/22', 1500), ('mnist/0', 1500), ('mnist/0', 30000)):
  if bsuite_id.startswith('mnist'):
    from bsuite_b200 import datasets
    os.environ[datasets.ENV_VAR] = datasets.write_synthetic_mnist('/tmp/bsb_sanitize_mnist', 256, 16, 0)
  results = []
  for bulk in ('1', '0'):
    os.environ['BSB_DEEP_SEA_BULK'] = bulk
    os.environ['BSB_EMIT_BULK'] = bulk
    env = bsuite_b200.load_from_id(bsuite_id, batch=batch, device='cuda', seed=3, track_episodes=True)
    acts = torch.as_tensor(env.random_actions(4, action_seed=1, first_step=0)).cuda()
    outs = [env.step(acts[i]).observation.clone() for i in range(4)]
    ts = env.rollout(3, action_seed=2)
    results.append(outs + [ts.observation.clone(), ts.reward.clone(), ts.step_type.clone()])
    torch.cuda.synchronize()
    env.close()
  same = all(torch.equal(a, b) for a, b in zip(*results))
  print(bsuite_id, batch, 'bulk == vector:', same, flush=True)
  assert same
# Graph-safe mode: device clock, chunk counter re-armed by the last CTA, the deterministic two-stage reduction.
for bsuite_id, batch in (('deep_sea/11', 20000), ('catch/0', 5000)):
  env = bsuite_b200.load_from_id(bsuite_id, batch=batch, device='cuda', seed=3, track_episodes=True)
  twin = bsuite_b200.load_from_id(bsuite_id, batch=batch, device='cuda', seed=3, track_episodes=True)
  graphed = env.capture(2, sample_actions=True, action_seed=5)
  for _ in range(3):
    got = graphed.replay()
    want = twin.rollout(2, action_seed=5)
    assert torch.equal(got.observation, want.observation) and torch.equal(got.reward, want.reward)
  env.step(torch.zeros(batch, dtype=torch.int32, device='cuda'))
  twin.step(torch.zeros(batch, dtype=torch.int32, device='cuda'))
  assert torch.equal(env.episode_stat_sums(), twin.episode_stat_sums()) and env.steps_done == twin.steps_done == 7
  print(bsuite_id, batch, 'graph replay == eager: True', flush=True)
  env.close(); twin.close()
# Host-driven steps: completion through the pinned mailbox, pre-launched doorbell kernels (under the sanitizer launches
# are serialised, so a pre-launched kernel times out and stands down: the cancel path), out-of-range actions.
for bsuite_id, batch in (('deep_sea/11', 20000), ('catch/0', 3000), ('mnist/0', 1500)):
  env = bsuite_b200.load_from_id(bsuite_id, batch=batch, device='cuda', seed=3, track_episodes=True)
  twin = bsuite_b200.load_from_id(bsuite_id, batch=batch, device='cuda', seed=3, track_episodes=True)
  host = env.make_host_buffers()
  acts = torch.as_tensor(env.random_actions(6, action_seed=1, first_step=0)).pin_memory()
  env.reset(); twin.reset()
  for t in range(6):
    got, obs = env.step_host(acts[t], host, prelaunch=(t >= 3))
    want = twin.step(acts[t].cuda())
    torch.cuda.synchronize()
    assert torch.equal(obs, want.observation) and torch.equal(got.reward, want.reward.cpu()) and torch.equal(got.step_type, want.step_type.cpu())
  bad = acts[0].cuda().clone(); bad[5] = 99
  env.step(bad); twin.step(bad.clamp(max=env.num_actions - 1))
  assert env.invalid_actions_seen() and torch.equal(env.episode_stat_sums(), twin.episode_stat_sums())
  print(bsuite_id, batch, 'host-driven steps == ordinary steps: True', flush=True)
  env.close(); twin.close()
split_steps()
# One-launch reduction over several environments and a whole lock-step in one graph.
from bsuite_b200 import suite
ids = ['catch/0', 'deep_sea/0', 'bandit_noise/0', 'cartpole/0', 'mnist/0', 'umbrella_length/0']
a, b = suite.SweepBatch(ids, lanes=300, device='cuda', seed=1), suite.SweepBatch(ids, lanes=300, device='cuda', seed=1, ring=2)
graphed = a.capture(1, lock_steps=2)
for _ in range(3):
  got, want = graphed.replay(), [b.rollout(1), b.rollout(1)]
  torch.cuda.synchronize()
  assert all(torch.equal(got[i][k].observation, want[i][k].observation) for i in range(2) for k in ids)
assert torch.equal(a.gather_returns(), b.local_returns().unsqueeze(0))
print('sweep graph == eager, one-launch reduction == per-id reductions: True', flush=True)
a.close(); b.close()
print('sanitize workload finished')
