"""The built-in `dm_env` stand-in (dm_env is not installed here).  What the reference relies on:
StepType 0/1/2, TimeStep namedtuple + helpers, restart / transition / termination / truncation
(utils/gym_wrapper.py:171), specs.Array / BoundedArray / DiscreteArray with validate() -- the check
EnvironmentTestMixin applies to every observation -- and Environment's default reward_spec / discount_spec
(gym_wrapper.py:93)."""

import numpy as np
import pytest

from bsuite_b200 import dm_env_compat as dm_env

specs = dm_env.specs


def test_step_types_and_timestep_helpers():
  assert [int(t) for t in dm_env.StepType] == [0, 1, 2]
  assert dm_env.StepType.FIRST.first() and dm_env.StepType.MID.mid() and dm_env.StepType.LAST.last()
  first = dm_env.restart(np.zeros(2))
  assert first.first() and first.reward is None and first.discount is None
  mid = dm_env.transition(1.5, np.ones(2))
  assert mid.mid() and mid.reward == 1.5 and mid.discount == 1.0
  assert dm_env.transition(0.0, None, discount=0.9).discount == 0.9
  last = dm_env.termination(-1.0, np.ones(2))
  assert last.last() and last.discount == 0.0
  cut = dm_env.truncation(2.0, np.ones(2))
  assert cut.last() and cut.discount == 1.0 and dm_env.truncation(2.0, None, 0.5).discount == 0.5
  assert mid._replace(reward=3.0).reward == 3.0                     # wrappers.py:166,171 use _replace
  step_type, reward, discount, observation = mid                    # it unpacks like the namedtuple it is
  assert (step_type, reward, discount) == (dm_env.StepType.MID, 1.5, 1.0) and observation.shape == (2,)
  # integer step types (the batched face stores int32) compare equal to the enum
  assert dm_env.TimeStep(np.int32(2), 0.0, 0.0, None).last()


def test_array_spec_validates_shape_and_dtype():
  spec = specs.Array((2, 3), np.float32, name='observation')
  assert spec.shape == (2, 3) and spec.dtype == np.float32 and spec.name == 'observation'
  assert spec.validate(np.zeros((2, 3), np.float32)).shape == (2, 3)
  with pytest.raises(ValueError, match='shape'):
    spec.validate(np.zeros((3, 2), np.float32))
  with pytest.raises(ValueError, match='dtype'):
    spec.validate(np.zeros((2, 3), np.float64))
  value = spec.generate_value()
  assert value.shape == (2, 3) and value.dtype == np.float32 and not value.any()
  assert spec == specs.Array((2, 3), np.float32, name='observation') and spec != specs.Array((2, 3), np.float64)
  assert hash(spec) == hash(specs.Array((2, 3), np.float32, name='observation'))
  assert spec.replace(name='other').name == 'other' and spec.replace(shape=(1,)).shape == (1,)
  assert 'Array(shape=(2, 3)' in repr(spec)


def test_bounded_array_spec():
  spec = specs.BoundedArray((2,), np.float32, minimum=0.0, maximum=[1.0, 2.0], name='board')
  spec.validate(np.array([1.0, 2.0], np.float32))
  with pytest.raises(ValueError, match='out of bounds'):
    spec.validate(np.array([1.5, 0.0], np.float32))
  with pytest.raises(ValueError, match='out of bounds'):
    spec.validate(np.array([-0.1, 0.0], np.float32))
  with pytest.raises(ValueError):
    specs.BoundedArray((2,), np.float32, minimum=1.0, maximum=0.0)
  with pytest.raises(ValueError):
    specs.BoundedArray((2,), np.float32, minimum=[0.0, 0.0, 0.0], maximum=1.0)     # not broadcastable
  with pytest.raises(ValueError):
    spec.minimum[...] = 5.0                                                        # read-only bounds
  assert spec.generate_value().tolist() == [0.0, 0.0]
  assert spec == specs.BoundedArray((2,), np.float32, 0.0, [1.0, 2.0], name='board')
  assert spec != specs.BoundedArray((2,), np.float32, 0.0, [1.0, 3.0], name='board')
  assert spec.replace(maximum=5.0).maximum == 5.0 and 'BoundedArray' in repr(spec)


def test_discrete_array_spec():
  spec = specs.DiscreteArray(3, name='action')
  assert spec.num_values == 3 and spec.shape == () and spec.dtype == np.int32
  assert spec.minimum == 0 and spec.maximum == 2
  spec.validate(np.int32(2))
  with pytest.raises(ValueError):
    spec.validate(np.int32(3))
  with pytest.raises(ValueError):
    spec.validate(2)                      # a Python int is int64: dtype mismatch, as in dm_env
  for bad in (0, -1, 2.5):
    with pytest.raises(ValueError):
      specs.DiscreteArray(bad)
  with pytest.raises(ValueError):
    specs.DiscreteArray(3, dtype=np.float32)
  assert specs.DiscreteArray(5, dtype=np.int64).dtype == np.int64
  assert spec.replace(num_values=7).num_values == 7 and 'num_values=3' in repr(spec)


def test_environment_base_class_defaults():
  class Tiny(dm_env.Environment):
    closed = False

    def reset(self):
      return dm_env.restart(np.zeros(1, np.float32))

    def step(self, action):
      return dm_env.termination(1.0, np.zeros(1, np.float32))

    def observation_spec(self):
      return specs.Array((1,), np.float32)

    def action_spec(self):
      return specs.DiscreteArray(2)

    def close(self):
      self.closed = True

  with pytest.raises(TypeError):
    dm_env.Environment()                  # abstract
  with Tiny() as env:
    assert env.reward_spec() == specs.Array((), float, name='reward')
    discount = env.discount_spec()
    assert (discount.minimum, discount.maximum, discount.name) == (0.0, 1.0, 'discount')
    env.reward_spec().validate(np.float64(env.step(0).reward))
  assert env.closed
