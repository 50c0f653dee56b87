"""A small, self-written stand-in for the `dm_env` package.

The reference's L0 substrate is `dm_env` (TimeStep / StepType / specs /
Environment ABC).  It is not installed in this
image and carries no arithmetic of the hot path, so `bsuite_b200` ships a
compatible module and prefers the real package whenever it is importable
(`bsuite_b200.dm_env` resolves to one or the other).

Only behaviour the reference relies on is provided:
  * `StepType` IntEnum FIRST/MID/LAST = 0/1/2 with `first()/mid()/last()`;
  * `TimeStep(step_type, reward, discount, observation)` namedtuple with the
    same predicates (the reference also calls `._replace`, wrappers.py:166);
  * `restart / transition / termination / truncation` constructors -- FIRST
    carries reward=None, discount=None; MID discount 1.0; LAST discount 0.0;
  * `Environment` ABC with default `reward_spec()/discount_spec()` (read by
    utils/gym_wrapper.py:93) and context-manager `close()`;
  * `specs.Array / BoundedArray / DiscreteArray`.
"""

import abc
import enum
from typing import Any, NamedTuple

import numpy as np


class StepType(enum.IntEnum):
  FIRST = 0
  MID = 1
  LAST = 2

  def first(self) -> bool:
    return self is StepType.FIRST

  def mid(self) -> bool:
    return self is StepType.MID

  def last(self) -> bool:
    return self is StepType.LAST


class TimeStep(NamedTuple):
  step_type: Any
  reward: Any
  discount: Any
  observation: Any

  def first(self) -> bool:
    return self.step_type == StepType.FIRST

  def mid(self) -> bool:
    return self.step_type == StepType.MID

  def last(self) -> bool:
    return self.step_type == StepType.LAST


def restart(observation):
  return TimeStep(StepType.FIRST, None, None, observation)


def transition(reward, observation, discount=1.0):
  return TimeStep(StepType.MID, reward, discount, observation)


def termination(reward, observation):
  return TimeStep(StepType.LAST, reward, 0.0, observation)


def truncation(reward, observation, discount=1.0):
  return TimeStep(StepType.LAST, reward, discount, observation)


class _Specs:
  """Namespace object exposed as `dm_env.specs`."""

  class Array:
    __slots__ = ('_shape', '_dtype', '_name')

    def __init__(self, shape, dtype, name=None):
      self._shape = tuple(int(d) for d in shape)
      self._dtype = np.dtype(dtype)
      self._name = name

    shape = property(lambda self: self._shape)
    dtype = property(lambda self: self._dtype)
    name = property(lambda self: self._name)

    def __repr__(self):
      return f'Array(shape={self.shape!r}, dtype={self.dtype!r}, name={self.name!r})'

    def __eq__(self, other):
      return (type(other) is type(self) and self.shape == other.shape and
              self.dtype == other.dtype and self.name == other.name)

    def __hash__(self):
      return hash((type(self).__name__, self.shape, self.dtype.str, self.name))

    def _fail(self, message):
      raise ValueError(f'{message} (spec {self!r})')

    def validate(self, value):
      value = np.asarray(value)
      if value.shape != self.shape:
        self._fail(f'Expected shape {self.shape} but found {value.shape}')
      if value.dtype != self.dtype:
        self._fail(f'Expected dtype {self.dtype} but found {value.dtype}')
      return value

    def generate_value(self):
      return np.zeros(self.shape, self.dtype)

    def replace(self, **kwargs):
      fields = dict(shape=self.shape, dtype=self.dtype, name=self.name)
      fields.update(kwargs)
      return type(self)(**fields)

  class BoundedArray(Array):
    __slots__ = ('_minimum', '_maximum')

    def __init__(self, shape, dtype, minimum, maximum, name=None):
      super().__init__(shape, dtype, name)
      lo = np.array(minimum, dtype=self.dtype)
      hi = np.array(maximum, dtype=self.dtype)
      np.broadcast(lo, np.empty(self.shape))  # raises on incompatible shapes
      np.broadcast(hi, np.empty(self.shape))
      if np.any(lo > hi):
        raise ValueError(f'minimum {lo} exceeds maximum {hi}')
      lo.setflags(write=False)
      hi.setflags(write=False)
      self._minimum, self._maximum = lo, hi

    minimum = property(lambda self: self._minimum)
    maximum = property(lambda self: self._maximum)

    def __repr__(self):
      return (f'BoundedArray(shape={self.shape!r}, dtype={self.dtype!r}, name={self.name!r}, '
              f'minimum={self.minimum}, maximum={self.maximum})')

    def __eq__(self, other):
      return (super().__eq__(other) and np.array_equal(self.minimum, other.minimum) and
              np.array_equal(self.maximum, other.maximum))

    __hash__ = None

    def validate(self, value):
      value = super().validate(value)
      if np.any(value < self.minimum) or np.any(value > self.maximum):
        self._fail(f'Value {value} is out of bounds [{self.minimum}, {self.maximum}]')
      return value

    def generate_value(self):
      return (np.ones(self.shape, self.dtype) * self.dtype.type(self.minimum)).astype(self.dtype)

    def replace(self, **kwargs):
      fields = dict(shape=self.shape, dtype=self.dtype, minimum=self.minimum,
                    maximum=self.maximum, name=self.name)
      fields.update(kwargs)
      return type(self)(**fields)

  class DiscreteArray(BoundedArray):
    __slots__ = ('_num_values',)

    def __init__(self, num_values, dtype=np.int32, name=None):
      if num_values <= 0 or not np.issubdtype(type(num_values), np.integer):
        raise ValueError(f'num_values must be a positive integer, got {num_values!r}')
      if not np.issubdtype(dtype, np.integer):
        raise ValueError(f'dtype must be integral, got {dtype!r}')
      super().__init__(shape=(), dtype=dtype, minimum=0, maximum=num_values - 1, name=name)
      self._num_values = int(num_values)

    num_values = property(lambda self: self._num_values)

    def __repr__(self):
      return f'DiscreteArray(shape=(), dtype={self.dtype!r}, name={self.name!r}, num_values={self.num_values})'

    def replace(self, **kwargs):
      fields = dict(num_values=self.num_values, dtype=self.dtype, name=self.name)
      fields.update(kwargs)
      return type(self)(**fields)


specs = _Specs


class Environment(abc.ABC):
  """Abstract RL environment (reset / step / observation_spec / action_spec)."""

  @abc.abstractmethod
  def reset(self) -> TimeStep:
    ...

  @abc.abstractmethod
  def step(self, action) -> TimeStep:
    ...

  @abc.abstractmethod
  def observation_spec(self):
    ...

  @abc.abstractmethod
  def action_spec(self):
    ...

  def reward_spec(self):
    return specs.Array(shape=(), dtype=float, name='reward')

  def discount_spec(self):
    return specs.BoundedArray(shape=(), dtype=float, minimum=0., maximum=1., name='discount')

  def close(self):
    pass

  def __enter__(self):
    return self

  def __exit__(self, exc_type, exc_value, traceback):
    self.close()
