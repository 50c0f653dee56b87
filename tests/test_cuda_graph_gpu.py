"""CUDA-graph capture of the transition kernel (-m gpu).

A graph freezes launch arguments, so the handle moves its step counter and chunk scheduler to device memory when a
launch is captured (include/bsuite_b200.h, "CUDA graphs").  Every test replays a graph several times and demands
the SAME results as an uncaptured twin environment fed the same actions: bit-exact outputs, accumulators, Logging
columns and step count -- then keeps going eagerly and through a state snapshot."""

import numpy as np
import pytest
import torch

import bsuite_b200

pytestmark = pytest.mark.gpu

CASES = [
    # deep_sea N = 32, 20 000 lanes: 625 chunks > the 444 resident warps -> the persistent grid's chunk counter is live
    ('deep_sea', dict(size=32, mapping_seed=42), 20000, 2),
    ('deep_sea', dict(size=10, deterministic=False, mapping_seed=3), 1000, 3),
    ('catch', dict(), 4099, 3),
    ('cartpole', dict(), 3000, 3),
    ('umbrella_chain', dict(chain_length=5, n_distractor=20), 2048, 3),
    ('memory_chain', dict(memory_length=4, num_bits=3), 777, 3),
    ('mnist', dict(), 300, 2),
    # 2 188 CTAs: above the grid size up to which the last CTA advances the device clock (a separate kernel does)
    ('bandit', dict(mapping_seed=1), 140000, 3),
    ('deep_sea', dict(size=6, mapping_seed=9), 80000, 2),    # persistent grid of 2 368 CTAs + external clock kernel
]


def _make(env_class, kwargs, batch, **extra):
  return bsuite_b200.make(env_class, batch=batch, device='cuda', seed=11,
                          engine_kwargs=dict(reward_dtype='float64', track_episodes=True, **extra), **kwargs)


def _same(ts_a, ts_b):
  for name in ('observation', 'reward', 'discount', 'step_type'):
    assert torch.equal(getattr(ts_a, name), getattr(ts_b, name)), name


def _same_books(a, b):
  info_a, info_b = a.bsuite_info(), b.bsuite_info()
  for key in info_a:
    assert torch.equal(info_a[key], info_b[key]), key
  stats_a, stats_b = a.episode_stats(), b.episode_stats()
  for key in stats_a:
    assert torch.equal(stats_a[key], stats_b[key]), key
  assert torch.equal(a.episode_stat_sums(), b.episode_stat_sums())
  assert a.steps_done == b.steps_done


@pytest.mark.parametrize('env_class,kwargs,batch,T', CASES)
@pytest.mark.parametrize('fused', [False, True], ids=['per_step', 'fused'])
def test_replayed_graph_equals_eager_steps(env_class, kwargs, batch, T, fused, mnist_dir):
  graphed_env, eager_env = _make(env_class, kwargs, batch), _make(env_class, kwargs, batch)
  n_actions = graphed_env.num_actions
  initial_state = eager_env.state_dict()
  gen = torch.Generator(device='cuda').manual_seed(5)
  # a few eager steps first: the capture must pick up a non-zero step count
  pre = torch.randint(0, n_actions, (3, batch), generator=gen, device='cuda', dtype=torch.int32)
  _same(graphed_env.rollout(3, actions=pre), eager_env.rollout(3, actions=pre))

  graphed = graphed_env.capture(T, fused=fused)
  for _ in range(5):
    acts = torch.randint(0, n_actions, (T, batch), generator=gen, device='cuda', dtype=torch.int32)
    graphed.actions.copy_(acts)
    _same(graphed.replay(), eager_env.rollout(T, actions=acts))
  assert graphed_env.steps_done == 3 + 5 * T
  _same_books(graphed_env, eager_env)

  # eager calls after the capture (device clock + programmatic dependent launch), interleaved with replays
  for k in range(4):
    acts = torch.randint(0, n_actions, (batch,), generator=gen, device='cuda', dtype=torch.int32)
    _same(graphed_env.step(acts), eager_env.step(acts))
  acts = torch.randint(0, n_actions, (T, batch), generator=gen, device='cuda', dtype=torch.int32)
  graphed.actions.copy_(acts)
  _same(graphed.replay(), eager_env.rollout(T, actions=acts))
  _same_books(graphed_env, eager_env)

  # snapshot of a graph-safe handle -> fresh (host-counted) handle, and back into the graph-safe one
  state = graphed_env.state_dict()
  fresh = _make(env_class, kwargs, batch)
  fresh.load_state_dict(state)
  graphed_env.load_state_dict(eager_env.state_dict())
  acts = torch.randint(0, n_actions, (T, batch), generator=gen, device='cuda', dtype=torch.int32)
  graphed.actions.copy_(acts)
  want = eager_env.rollout(T, actions=acts)
  _same(fresh.rollout(T, actions=acts), want)
  _same(graphed.replay(), want)
  _same_books(fresh, eager_env)
  _same_books(graphed_env, eager_env)

  # a snapshot OLDER than the capture: the device clock then holds a negative offset from the captured base
  graphed_env.load_state_dict(initial_state)
  eager_env.load_state_dict(initial_state)
  assert graphed_env.steps_done == eager_env.steps_done == 0
  for _ in range(3):
    acts = torch.randint(0, n_actions, (T, batch), generator=gen, device='cuda', dtype=torch.int32)
    graphed.actions.copy_(acts)
    _same(graphed.replay(), eager_env.rollout(T, actions=acts))
  _same_books(graphed_env, eager_env)


@pytest.mark.parametrize('env_class,kwargs,batch', [
    ('catch', dict(), 5000),
    ('deep_sea', dict(size=32, mapping_seed=42), 20000),
    ('mountain_car', dict(), 2500),
])
@pytest.mark.parametrize('fused', [False, True], ids=['per_step', 'fused'])
def test_graph_with_device_sampled_actions_advances_the_action_stream(env_class, kwargs, batch, fused):
  """Actions sampled on the device are keyed by the step index: replays must continue the stream (a frozen step
  argument would repeat the first replay's actions), and equal one uncaptured rollout of the same total length."""
  T, R, seed = 4, 4, 77
  graphed_env, eager_env = _make(env_class, kwargs, batch), _make(env_class, kwargs, batch)
  graphed = graphed_env.capture(T, sample_actions=True, fused=fused, action_seed=seed)
  eager_out = eager_env.make_buffers(T * R, with_actions=True)
  want = eager_env.rollout(T * R, action_seed=seed, out=eager_out)
  mirror = torch.as_tensor(eager_env.random_actions(T * R, action_seed=seed, first_step=0), device='cuda')
  assert torch.equal(eager_out.actions, mirror)
  for r in range(R):
    got = graphed.replay()
    assert torch.equal(graphed.buffers.actions, eager_out.actions[r * T:(r + 1) * T])
    for name in ('observation', 'reward', 'discount', 'step_type'):
      assert torch.equal(getattr(got, name), getattr(want, name)[r * T:(r + 1) * T]), name
  _same_books(graphed_env, eager_env)


def test_log_point_inside_a_graph():
  """The device-side reduction of the Logging columns captured together with the steps it summarises."""
  batch, T = 4096, 8
  graphed_env, eager_env = _make('catch', dict(), batch), _make('catch', dict(), batch)
  out = graphed_env.make_buffers(T, with_actions=True)
  graphed_env.rollout(T, action_seed=3, out=out)            # eager pass: loads the module outside the capture
  eager_env.rollout(T, action_seed=3)
  torch.cuda.synchronize()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    graphed_env.rollout(T, action_seed=3, out=out)
    sums = graphed_env.episode_stat_sums()
  for _ in range(3):
    graph.replay()
    eager_env.rollout(T, action_seed=3)
    assert torch.equal(sums, eager_env.episode_stat_sums())
  assert float(sums[0]) > 0


def test_sweep_lock_step_captured_in_one_graph_equals_eager_rollouts(mnist_dir):
  """SweepBatch.capture: every id's launch in ONE graph (fork / join over the ids' streams); replays must equal an
  uncaptured twin bit for bit, and mix with eager rollouts."""
  from bsuite_b200 import suite
  ids = ['catch/0', 'deep_sea/0', 'bandit_noise/0', 'cartpole/0', 'mnist/0', 'umbrella_length/0', 'memory_size/0']
  a = suite.SweepBatch(ids, lanes=512, device='cuda', seed=1)
  b = suite.SweepBatch(ids, lanes=512, device='cuda', seed=1)
  graphed = a.capture(num_steps=1)
  for round_ in range(40):
    got = graphed.replay()
    want = b.rollout(1)
    torch.cuda.synchronize()
    for k in ids:
      for field in ('step_type', 'reward', 'discount', 'observation'):
        assert torch.equal(getattr(got[k], field), getattr(want[k], field)), (round_, k, field)
    if round_ == 17:                     # an eager rollout in between
      a.rollout(3); b.rollout(3)
  assert torch.equal(a.gather_returns(), b.gather_returns())
  a.close(); b.close()
