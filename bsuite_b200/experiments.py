"""Experiment name + kwargs -> engine configuration.

Each builder takes the keyword arguments of the reference constructor or `load`
factory it stands for (same names, same defaults) and returns an `EnvSpec`: the
flattened `bsb_config` fields plus the host tables, which are produced with the
SAME numpy calls the reference constructors make, so they are equal by
construction (deep_sea action mapping, bandit arm permutation, discounting
bonus arm).
"""

import dataclasses
from typing import Any, Callable, Dict, Mapping, Optional, Tuple

import numpy as np

from bsuite_b200 import _lib
from bsuite_b200 import datasets


@dataclasses.dataclass
class EnvSpec:
  family: int
  fields: Dict[str, Any]                     # bsb_config scalar fields
  obs_shape: Tuple[int, ...]
  num_actions: int
  bsuite_num_episodes: int
  seed: Optional[int] = None                 # the reference's `seed` kwarg
  wrapper: int = _lib.WRAP_NONE
  table: Optional[np.ndarray] = None
  table2: Optional[np.ndarray] = None
  obs_spec_name: str = 'observation'
  obs_bounds: Optional[Tuple[float, float]] = None   # BoundedArray(min, max) when set
  action_dtype: Any = np.int32


# ---- environment classes ---------------------------------------------------
def deep_sea(size: int, deterministic: bool = True, unscaled_move_cost: float = 0.01,
             randomize_actions: bool = True, seed: Optional[int] = None,
             mapping_seed: Optional[int] = None) -> EnvSpec:
  """environments/deep_sea.py:51-101."""
  if randomize_actions:
    mapping = np.random.RandomState(mapping_seed).binomial(1, 0.5, [size, size])   # deep_sea.py:80-81
  else:
    mapping = np.ones([size, size])                                               # deep_sea.py:85
  return EnvSpec(
      family=_lib.DEEP_SEA,
      fields=dict(size=int(size), deterministic=int(bool(deterministic)),
                  unscaled_move_cost=float(unscaled_move_cost)),
      table=np.ascontiguousarray(mapping, dtype=np.uint8),
      obs_shape=(int(size), int(size)), num_actions=2, bsuite_num_episodes=10000, seed=seed)


def catch(rows: int = 10, columns: int = 5, seed: Optional[int] = None) -> EnvSpec:
  """environments/catch.py:45-66; BoundedArray[0, 1] observation (:99-102)."""
  return EnvSpec(family=_lib.CATCH, fields=dict(rows=int(rows), columns=int(columns)),
                 obs_shape=(int(rows), int(columns)), num_actions=3, bsuite_num_episodes=10000,
                 seed=seed, obs_bounds=(0., 1.), action_dtype=np.int64)


def cartpole(height_threshold: float = 0.8, x_threshold: float = 3., timescale: float = 0.01,
             max_time: float = 10., init_range: float = 0.05, seed: Optional[int] = None) -> EnvSpec:
  """environments/cartpole.py:81-115."""
  return EnvSpec(
      family=_lib.CARTPOLE,
      fields=dict(height_threshold=float(height_threshold), x_threshold=float(x_threshold),
                  timescale=float(timescale), max_time=float(max_time), init_range=float(init_range)),
      obs_shape=(1, 6), num_actions=3, bsuite_num_episodes=1000, seed=seed, action_dtype=np.int64)


def cartpole_swingup(height_threshold: float = 0.5, theta_dot_threshold: float = 1.,
                     x_reward_threshold: float = 1., move_cost: float = 0.1, x_threshold: float = 3.,
                     timescale: float = 0.01, max_time: float = 10., init_range: float = 0.05,
                     seed: Optional[int] = None) -> EnvSpec:
  """experiments/cartpole_swingup/cartpole_swingup.py:41-79; obs spec is named 'state' (:134)."""
  return EnvSpec(
      family=_lib.CARTPOLE_SWINGUP,
      fields=dict(height_threshold=float(height_threshold), theta_dot_threshold=float(theta_dot_threshold),
                  x_reward_threshold=float(x_reward_threshold), move_cost=float(move_cost),
                  x_threshold=float(x_threshold), timescale=float(timescale), max_time=float(max_time),
                  init_range=float(init_range)),
      obs_shape=(1, 8), num_actions=3, bsuite_num_episodes=1000, seed=seed, obs_spec_name='state',
      action_dtype=np.int64)


def mountain_car(max_steps: int = 1000, seed: Optional[int] = None) -> EnvSpec:
  """environments/mountain_car.py:36-60."""
  return EnvSpec(family=_lib.MOUNTAIN_CAR, fields=dict(max_steps=int(max_steps)), obs_shape=(1, 3),
                 num_actions=3, bsuite_num_episodes=1000, seed=seed)


def memory_chain(memory_length: int, num_bits: int = 1, seed: Optional[int] = None) -> EnvSpec:
  """environments/memory_chain.py:37-58."""
  return EnvSpec(family=_lib.MEMORY_CHAIN,
                 fields=dict(memory_length=int(memory_length), num_bits=int(num_bits)),
                 obs_shape=(1, int(num_bits) + 2), num_actions=2, bsuite_num_episodes=10000, seed=seed)


def bandit(mapping_seed: Optional[int] = None, num_actions: int = 11) -> EnvSpec:
  """environments/bandit.py:35-51."""
  rng = np.random.RandomState(mapping_seed)
  order = rng.choice(range(num_actions), size=num_actions, replace=False)        # bandit.py:45-46
  rewards = np.linspace(0, 1, num_actions)[order]                                # bandit.py:47
  return EnvSpec(family=_lib.BANDIT, fields=dict(num_actions=int(num_actions)),
                 table=np.ascontiguousarray(rewards, dtype=np.float64), obs_shape=(1, 1),
                 num_actions=int(num_actions), bsuite_num_episodes=10000)


def umbrella_chain(chain_length: int, n_distractor: int = 0, seed: Optional[int] = None) -> EnvSpec:
  """environments/umbrella_chain.py:39-58."""
  return EnvSpec(family=_lib.UMBRELLA_CHAIN,
                 fields=dict(chain_length=int(chain_length), n_distractor=int(n_distractor)),
                 obs_shape=(1, 3 + int(n_distractor)), num_actions=2, bsuite_num_episodes=10000, seed=seed)


def discounting_chain(mapping_seed: Optional[int] = None) -> EnvSpec:
  """environments/discounting_chain.py:40-61."""
  if mapping_seed is None:
    mapping_seed = np.random.randint(0, 5)                                       # :49-50
  else:
    mapping_seed = mapping_seed % 5                                              # :52
  rewards = np.ones(5)
  rewards[mapping_seed] += 0.1                                                   # :55-56
  return EnvSpec(family=_lib.DISCOUNTING_CHAIN, fields={}, table=rewards, obs_shape=(1, 2),
                 num_actions=5, bsuite_num_episodes=1000)


def mnist(fraction: float = 1., seed: Optional[int] = None, data_dir: Optional[str] = None) -> EnvSpec:
  """environments/mnist.py:36-59 (dataset: utils/datasets.py:42-69, parsed as int8)."""
  images, labels = datasets.load_mnist_train(data_dir)
  num_data = int(fraction * len(labels))                                         # mnist.py:46-48
  images = np.ascontiguousarray(images[:num_data])
  labels = np.ascontiguousarray(labels[:num_data])
  return EnvSpec(family=_lib.MNIST,
                 fields=dict(num_data=num_data, image_rows=int(images.shape[1]), image_cols=int(images.shape[2])),
                 table=images, table2=labels, obs_shape=tuple(int(d) for d in images.shape[1:]),
                 num_actions=10, bsuite_num_episodes=10000, seed=seed)


# ---- reward wrappers (utils/wrappers.py:250-373) ----------------------------
def _with_noise(spec: EnvSpec, noise_scale: float, num_episodes: int) -> EnvSpec:
  spec.wrapper = _lib.WRAP_REWARD_NOISE
  spec.fields['noise_scale'] = float(noise_scale)
  spec.bsuite_num_episodes = num_episodes
  return spec


def _with_scale(spec: EnvSpec, reward_scale: float, num_episodes: int) -> EnvSpec:
  spec.wrapper = _lib.WRAP_REWARD_SCALE
  spec.fields['reward_scale'] = float(reward_scale)
  spec.bsuite_num_episodes = num_episodes
  return spec


# ---- experiment `load` factories (experiments/<name>/<name>.py) -------------
def _bandit_noise(noise_scale, seed, mapping_seed, num_actions=11):          # bandit_noise.py:27-34
  spec = _with_noise(bandit(mapping_seed, num_actions=num_actions), noise_scale, 10000)
  spec.seed = seed
  return spec


def _bandit_scale(reward_scale, seed, mapping_seed):                         # bandit_scale.py:27-34
  spec = _with_scale(bandit(mapping_seed=mapping_seed), reward_scale, 10000)
  spec.seed = seed
  return spec


def _deep_sea_stochastic(size: int, mapping_seed=0):                         # deep_sea_stochastic.py:22-30
  return deep_sea(size=size, deterministic=False, mapping_seed=mapping_seed)


def _memory_len(memory_length: int, seed: Optional[int] = 0):                # memory_len.py:31-39
  return memory_chain(memory_length=memory_length, num_bits=1, seed=seed)


def _memory_size(num_bits: int, seed: Optional[int] = 0):                    # memory_size.py:31-39
  return memory_chain(memory_length=2, num_bits=num_bits, seed=seed)


def _umbrella_distract(n_distractor: int, seed=0):                           # umbrella_distract.py:22-30
  return umbrella_chain(chain_length=20, n_distractor=n_distractor, seed=seed)


# experiment name -> builder; the keys and keyword arguments are those of
# bsuite.bsuite.EXPERIMENT_NAME_TO_ENVIRONMENT (bsuite/bsuite.py:57-81).
EXPERIMENT_NAME_TO_SPEC: Mapping[str, Callable[..., EnvSpec]] = dict(
    bandit=bandit,
    bandit_noise=_bandit_noise,
    bandit_scale=_bandit_scale,
    cartpole=cartpole,
    cartpole_noise=lambda noise_scale, seed: _with_noise(cartpole(seed=seed), noise_scale, 1000),
    cartpole_scale=lambda reward_scale, seed: _with_scale(cartpole(seed=seed), reward_scale, 1000),
    cartpole_swingup=cartpole_swingup,
    catch=catch,
    catch_noise=lambda noise_scale, seed: _with_noise(catch(seed=seed), noise_scale, 10000),
    catch_scale=lambda reward_scale, seed: _with_scale(catch(seed=seed), reward_scale, 10000),
    deep_sea=deep_sea,
    deep_sea_stochastic=_deep_sea_stochastic,
    discounting_chain=discounting_chain,
    memory_len=_memory_len,
    memory_size=_memory_size,
    mnist=mnist,
    mnist_noise=lambda noise_scale, seed: _with_noise(mnist(seed=seed), noise_scale, 10000),
    mnist_scale=lambda reward_scale, seed: _with_scale(mnist(seed=seed), reward_scale, 10000),
    mountain_car=mountain_car,
    mountain_car_noise=lambda noise_scale, seed: _with_noise(mountain_car(seed=seed), noise_scale, 1000),
    mountain_car_scale=lambda reward_scale, seed: _with_scale(mountain_car(seed=seed), reward_scale, 1000),
    umbrella_distract=_umbrella_distract,
    umbrella_length=umbrella_chain,
)

# Raw environment classes, for callers that construct environments directly
# (the reference's tests do: deep_sea_test.py:27-31 etc.).
ENVIRONMENT_CLASSES: Mapping[str, Callable[..., EnvSpec]] = dict(
    deep_sea=deep_sea, catch=catch, cartpole=cartpole, cartpole_swingup=cartpole_swingup,
    mountain_car=mountain_car, memory_chain=memory_chain, bandit=bandit, umbrella_chain=umbrella_chain,
    discounting_chain=discounting_chain, mnist=mnist)
