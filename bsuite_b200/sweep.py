"""The bsuite experiment registry: 23 experiments, 468 `bsuite_id`s.

Mirrors the public names of the reference's `bsuite/sweep.py:134-150`
(`SETTINGS`, `SWEEP`, `TAGS`, `TESTING`, `EPISODES`, `SEPARATOR`, and one tuple of
ids per experiment such as `DEEP_SEA`) so that code written against
`bsuite.sweep` finds the same ids mapped to the same keyword arguments.  The
tables are data restated from `bsuite/experiments/<name>/sweep.py` (one citation
per entry below); they are generated here from compact rules rather than
spread over 23 modules.
"""

import types
from typing import Any, Dict, Mapping, Tuple

SEPARATOR = '/'
IGNORE_FOR_TESTING = ('_noise', '_scale')

_NOISE_SCALES = (0.1, 0.3, 1.0, 3., 10.)        # e.g. catch_noise/sweep.py:22-27
_REWARD_SCALES = (0.001, 0.03, 1.0, 30., 1000.)  # e.g. catch_scale/sweep.py:22-27
# memory_len/sweep.py:22-26, umbrella_*/sweep.py: 1..10, 12, 14, 17, 20, 25, 30..100 step 10
_LOG_SPACED_100 = tuple(range(1, 11)) + (12, 14, 17, 20, 25) + tuple(range(30, 105, 10))
# memory_size/sweep.py:23-27: ... 30, 40
_LOG_SPACED_40 = tuple(range(1, 11)) + (12, 14, 17, 20, 25) + tuple(range(30, 50, 10))


def _wrapped(key, scales, inner):
  """Five scales x four replicas, replica-major inside each scale."""
  return tuple(dict({key: scale}, **inner(n)) for scale in scales for n in range(4))


def _experiments():
  """(name, settings, num_episodes, tags) in the reference's registration order."""
  seedless = lambda n: {'seed': None}
  bandit_inner = lambda n: {'seed': None, 'mapping_seed': n}
  deep_sea_sizes = tuple({'size': n, 'mapping_seed': 42} for n in range(10, 51, 2))
  return (
      # bandit/sweep.py:20-24
      ('bandit', tuple({'mapping_seed': n} for n in range(20)), 10000, ('basic',)),
      ('bandit_noise', _wrapped('noise_scale', _NOISE_SCALES, bandit_inner), 10000, ('noise',)),
      ('bandit_scale', _wrapped('reward_scale', _REWARD_SCALES, bandit_inner), 10000, ('scale',)),
      # cartpole/sweep.py:20-22
      ('cartpole', tuple({'seed': None} for _ in range(20)), 1000,
       ('basic', 'credit_assignment', 'generalization')),
      ('cartpole_noise', _wrapped('noise_scale', _NOISE_SCALES, seedless), 1000, ('noise', 'generalization')),
      ('cartpole_scale', _wrapped('reward_scale', _REWARD_SCALES, seedless), 1000, ('scale', 'generalization')),
      # cartpole_swingup/sweep.py:22-25
      ('cartpole_swingup',
       tuple({'height_threshold': n / 20, 'x_reward_threshold': 1 - n / 20} for n in range(20)), 1000,
       ('exploration', 'generalization')),
      # catch/sweep.py:20-22
      ('catch', tuple({'seed': None} for _ in range(20)), 10000, ('basic', 'credit_assignment')),
      ('catch_noise', _wrapped('noise_scale', _NOISE_SCALES, seedless), 10000, ('noise', 'credit_assignment')),
      ('catch_scale', _wrapped('reward_scale', _REWARD_SCALES, seedless), 10000, ('scale', 'credit_assignment')),
      # deep_sea/sweep.py:20-22, deep_sea_stochastic/sweep.py:22-24
      ('deep_sea', deep_sea_sizes, 10000, ('exploration',)),
      ('deep_sea_stochastic', deep_sea_sizes, 10000, ('exploration', 'noise')),
      # discounting_chain/sweep.py:20-22
      ('discounting_chain', tuple({'mapping_seed': n} for n in range(20)), 1000, ('credit_assignment',)),
      # memory_len/sweep.py:20-29, memory_size/sweep.py:22-30
      ('memory_len', tuple({'memory_length': n} for n in _LOG_SPACED_100), 10000, ('memory',)),
      ('memory_size', tuple({'num_bits': n} for n in _LOG_SPACED_40), 10000, ('memory',)),
      # mnist/sweep.py:20-22
      ('mnist', tuple({'seed': None} for _ in range(20)), 10000, ('basic', 'generalization')),
      ('mnist_noise', _wrapped('noise_scale', _NOISE_SCALES, seedless), 10000, ('noise', 'generalization')),
      ('mnist_scale', _wrapped('reward_scale', _REWARD_SCALES, seedless), 10000, ('scale', 'generalization')),
      # mountain_car/sweep.py:20-22
      ('mountain_car', tuple({'seed': None} for _ in range(20)), 1000, ('basic', 'generalization')),
      ('mountain_car_noise', _wrapped('noise_scale', _NOISE_SCALES, seedless), 1000, ('noise', 'generalization')),
      ('mountain_car_scale', _wrapped('reward_scale', _REWARD_SCALES, seedless), 1000, ('scale', 'generalization')),
      # umbrella_distract/sweep.py:22-30, umbrella_length/sweep.py:20-28
      ('umbrella_distract', tuple({'n_distractor': n} for n in _LOG_SPACED_100), 10000,
       ('credit_assignment', 'noise')),
      ('umbrella_length', tuple({'chain_length': n, 'n_distractor': 20} for n in _LOG_SPACED_100), 10000,
       ('credit_assignment', 'noise')),
  )


def _freeze(d: Dict[str, Any]) -> Mapping[str, Any]:
  return types.MappingProxyType(dict(d))


_settings: Dict[str, Mapping[str, Any]] = {}
_sweep = []
_tags: Dict[str, list] = {}
_testing = []
_episodes: Dict[str, int] = {}
_by_experiment: Dict[str, Tuple[str, ...]] = {}

for _name, _exp_settings, _num_episodes, _exp_tags in _experiments():
  _ids = []
  for _i, _setting in enumerate(_exp_settings):
    _id = f'{_name}{SEPARATOR}{_i}'
    # sweep.py:86-94: setting 0 of every experiment not ending in _noise/_scale
    if _i == 0 and not _name.endswith(IGNORE_FOR_TESTING):
      _testing.append(_id)
    _ids.append(_id)
    _settings[_id] = _freeze(_setting)
    _episodes[_id] = _num_episodes
  for _tag in _exp_tags:
    _tags.setdefault(_tag, []).extend(_ids)
  _sweep.extend(_ids)
  _by_experiment[_name] = tuple(_ids)
  globals()[_name.upper()] = tuple(_ids)   # BANDIT, DEEP_SEA, ... (sweep.py:109-131)

# bsuite_id -> constructor kwargs (sweep.py:134-135)
SETTINGS: Mapping[str, Mapping[str, Any]] = types.MappingProxyType(_settings)
# every bsuite_id (sweep.py:138)
SWEEP: Tuple[str, ...] = tuple(_sweep)
# tag -> ids (sweep.py:143-144)
TAGS: Mapping[str, Tuple[str, ...]] = types.MappingProxyType({k: tuple(v) for k, v in _tags.items()})
# representative subset for tests (sweep.py:147)
TESTING: Tuple[str, ...] = tuple(_testing)
# bsuite_id -> bsuite_num_episodes (sweep.py:150)
EPISODES: Mapping[str, int] = types.MappingProxyType(_episodes)
# experiment name -> its ids
BY_EXPERIMENT: Mapping[str, Tuple[str, ...]] = types.MappingProxyType(_by_experiment)
