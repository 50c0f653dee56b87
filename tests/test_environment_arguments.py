"""Argument handling of the batched face on the explicit host path: what is converted, what is refused."""

import numpy as np
import pytest
import torch

import bsuite_b200


def _env(**kwargs):
  return bsuite_b200.load_from_id('catch/0', batch=6, device='cpu', seed=2, **kwargs)


def test_device_and_rng_arguments():
  with pytest.raises(ValueError, match='unsupported device'):
    bsuite_b200.load_from_id('catch/0', batch=2, device='meta')
  with pytest.raises(ValueError, match='rng must be'):
    _env(rng='pcg64')
  with pytest.raises(ValueError, match='Seed must be between 0 and 2\\*\\*32 - 1'):      # numpy's own message
    bsuite_b200.load_from_id('catch/0', batch=2, device='cpu', seed=2**32, rng='mt19937')
  big = bsuite_b200.load_from_id('catch/0', batch=2, device='cpu', seed=2**40)            # Philox keys are 64-bit
  assert big.seed == 2**40
  env = _env()
  assert (env.batch, env.device.type, env.lane_offset, env.num_actions, env.obs_shape) == (6, 'cpu', 0, 3, (10, 5))
  assert env.info_names == ('total_regret',)


def test_step_accepts_what_converts_to_int32_actions_and_refuses_other_shapes():
  reference = _env()
  want = [reference.step(torch.full((6,), a, dtype=torch.int32)) for a in (0, 2, 1)]
  for make in (lambda a: [a] * 6, lambda a: np.full(6, a, np.int64), lambda a: torch.full((6,), a, dtype=torch.int64),
               lambda a: torch.full((12,), a, dtype=torch.int32)[::2]):               # list, int64, non-contiguous
    env = _env()
    for ts, a in zip(want, (0, 2, 1)):
      got = env.step(make(a))
      assert torch.equal(got.observation, ts.observation) and torch.equal(got.step_type, ts.step_type)
  env = _env()
  for bad in (torch.zeros(5, dtype=torch.int32), torch.zeros((6, 1), dtype=torch.int32), [0, 1]):
    with pytest.raises(ValueError, match='actions must have shape'):
      env.step(bad)
  with pytest.raises(ValueError, match='actions must have shape'):
    env.rollout(4, actions=torch.zeros((3, 6), dtype=torch.int32))
  assert env.steps_done == 0                    # refused calls do not advance the environment


def test_step_host_arguments_and_float64_rewards():
  env = _env(reward_dtype='float64')
  twin = _env(reward_dtype='float64')
  host = env.make_host_buffers(with_observation=True)
  assert host.reward.dtype == torch.float64 and host.observation.shape == (6, 10, 5)
  mixed = env.make_mixed_buffers()              # host environment: ordinary buffers
  assert mixed.reward.dtype == torch.float64 and mixed.observation.device.type == 'cpu'
  actions = torch.tensor([0, 1, 2, 0, 1, 2], dtype=torch.int32)
  for _ in range(12):
    got, device_obs = env.step_host(actions, host)
    want = twin.step(actions)
    assert torch.equal(got.reward, want.reward) and torch.equal(got.step_type, want.step_type)
    assert torch.equal(host.observation, want.observation) and torch.equal(device_obs, want.observation)
  got, _ = env.step_host(np.array([0, 1, 2, 0, 1, 2], np.int32), host)          # a numpy int32 array is fine
  assert torch.equal(got.step_type, twin.step(actions).step_type)
  for bad in ([0, 1, 2, 0, 1, 2], torch.zeros(6, dtype=torch.int64), torch.zeros(5, dtype=torch.int32)):
    with pytest.raises(ValueError, match='step_host takes a CPU int32 tensor'):
      env.step_host(bad, host)


def test_accumulator_and_snapshot_errors():
  env = _env()
  with pytest.raises(RuntimeError, match='track_episodes=True'):
    env.episode_stats()
  other = bsuite_b200.load_from_id('catch/0', batch=7, device='cpu', seed=2)
  with pytest.raises(ValueError, match='different environment'):
    other.load_state_dict(env.state_dict())
  reseeded = bsuite_b200.load_from_id('catch/0', batch=6, device='cpu', seed=3)
  with pytest.raises(ValueError, match='seed, lane_offset'):
    reseeded.load_state_dict(env.state_dict())
  with pytest.raises(RuntimeError, match='CUDA graphs need a CUDA environment'):
    env.capture(2)
  env.close()
  env.close()                                   # idempotent
