"""Generates tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py            # regenerates every fixture
    python oracle/gen_golden.py --check    # regenerates in memory and diffs against the committed files

Each fixture is one case: an environment class + kwargs (+ reward wrapper),
8 lanes with global ids 0..7, an action matrix [T, 8], and for every call the
reference's (step_type, reward, discount, observation) per lane plus the final
bsuite_info().  Lane i's reference environment is constructed with
`seed = lane_seed(rng, seed, i)` (oracle/reference_runner.py): a per-lane
Philox bit generator, or for rng == 'mt19937' the plain integer seed + i, i.e.
the unpatched reference.  `reset_at` lists call indices where reset() is called
instead of step().

Also writes tests/golden/known_answers.json: the SURVEY.md 8c table (digest,
reward sum, #LAST, bsuite_info) for single reference environments.
"""

import argparse
import json
import os
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _ROOT)

from oracle import reference_runner as rr  # noqa: E402

GOLDEN_DIR = os.path.join(_ROOT, 'tests', 'golden')
NUM_LANES = 8
MNIST_SEED, MNIST_TRAIN, MNIST_TEST = 0, 256, 16


def case(name, env_class, kwargs, steps, rng='philox', seed=20240229, wrapper=None, wrapper_arg=0.0, reset_at=()):
  return dict(name=name, env_class=env_class, kwargs=kwargs, steps=steps, rng=rng, seed=seed, wrapper=wrapper,
              wrapper_arg=wrapper_arg, reset_at=list(reset_at))


CASES = [
    # ---- integer / grid families: bit-exact --------------------------------
    case('deep_sea_10', 'deep_sea', dict(size=10, mapping_seed=42), 200),
    case('deep_sea_32', 'deep_sea', dict(size=32, mapping_seed=42), 110),
    case('deep_sea_50_resets', 'deep_sea', dict(size=50, mapping_seed=42), 120, reset_at=(0, 17, 18, 70)),
    case('deep_sea_stochastic_10', 'deep_sea', dict(size=10, deterministic=False, mapping_seed=42), 400),
    case('deep_sea_stochastic_7_odd', 'deep_sea', dict(size=7, deterministic=False, mapping_seed=3), 150),
    case('deep_sea_debug_mapping', 'deep_sea', dict(size=6, randomize_actions=False), 60),
    case('catch', 'catch', dict(), 220),
    case('catch_resets', 'catch', dict(), 60, reset_at=(0, 5, 6, 30)),
    case('catch_7x3', 'catch', dict(rows=7, columns=3), 80),
    case('catch_noise_3', 'catch', dict(), 120, wrapper='noise', wrapper_arg=3.0),
    case('catch_scale_30', 'catch', dict(), 60, wrapper='scale', wrapper_arg=30.0),
    case('memory_len_6', 'memory_chain', dict(memory_length=6, num_bits=1), 200),
    case('memory_len_1', 'memory_chain', dict(memory_length=1, num_bits=1), 60),
    case('memory_size_40', 'memory_chain', dict(memory_length=2, num_bits=40), 100),
    case('memory_7_bits_3', 'memory_chain', dict(memory_length=7, num_bits=3), 120, reset_at=(0, 4, 50)),
    case('bandit_3', 'bandit', dict(mapping_seed=3), 60),
    case('bandit_noise_1', 'bandit', dict(mapping_seed=2), 80, wrapper='noise', wrapper_arg=1.0),
    case('bandit_scale_0p001', 'bandit', dict(mapping_seed=0), 40, wrapper='scale', wrapper_arg=0.001),
    case('umbrella_length_12', 'umbrella_chain', dict(chain_length=12, n_distractor=20), 150),
    case('umbrella_distract_100', 'umbrella_chain', dict(chain_length=20, n_distractor=100), 90),
    case('umbrella_chain_1', 'umbrella_chain', dict(chain_length=1, n_distractor=0), 40),
    case('umbrella_chain_2_resets', 'umbrella_chain', dict(chain_length=2, n_distractor=3), 50, reset_at=(3, 4)),
    case('discounting_chain_2', 'discounting_chain', dict(mapping_seed=2), 260),
    case('mnist', 'mnist', dict(), 40),
    case('mnist_half', 'mnist', dict(fraction=0.5), 24),
    case('mnist_noise_0p3', 'mnist', dict(), 24, wrapper='noise', wrapper_arg=0.3),
    # ---- float dynamics: <= 1e-6 ---------------------------------------------
    case('cartpole', 'cartpole', dict(), 700),
    case('cartpole_resets', 'cartpole', dict(), 120, reset_at=(0, 10, 60)),
    case('cartpole_noise_0p1', 'cartpole', dict(), 250, wrapper='noise', wrapper_arg=0.1),
    case('cartpole_scale_1000', 'cartpole', dict(), 150, wrapper='scale', wrapper_arg=1000.0),
    case('cartpole_swingup_5', 'cartpole_swingup', dict(height_threshold=5 / 20, x_reward_threshold=1 - 5 / 20), 1300),
    case('mountain_car', 'mountain_car', dict(), 1150),
    case('mountain_car_2', 'mountain_car', dict(max_steps=2), 40),
    case('mountain_car_noise_10', 'mountain_car', dict(max_steps=30), 100, wrapper='noise', wrapper_arg=10.0),
    # ---- MT19937: the UNPATCHED reference with integer seeds seed + lane ------
    case('mt_catch', 'catch', dict(), 150, rng='mt19937', seed=0),
    case('mt_deep_sea_stochastic_10', 'deep_sea', dict(size=10, deterministic=False, mapping_seed=42), 300,
         rng='mt19937', seed=0),
    case('mt_memory_size_40', 'memory_chain', dict(memory_length=2, num_bits=40), 400, rng='mt19937', seed=0),
    case('mt_umbrella_12_20', 'umbrella_chain', dict(chain_length=12, n_distractor=20), 700, rng='mt19937', seed=0),
    case('mt_cartpole', 'cartpole', dict(), 400, rng='mt19937', seed=0),
    case('mt_mountain_car', 'mountain_car', dict(max_steps=50), 160, rng='mt19937', seed=5),
    case('mt_catch_noise_1', 'catch', dict(), 700, rng='mt19937', seed=11, wrapper='noise', wrapper_arg=1.0),
    case('mt_mnist', 'mnist', dict(), 30, rng='mt19937', seed=4),
]

# Single reference environments, reset() then 1000 step() calls.
KNOWN_ANSWERS = [
    ('load_from_id', 'deep_sea/0'), ('load_from_id', 'deep_sea/11'), ('load_from_id', 'discounting_chain/0'),
    ('load_from_id', 'bandit/0'), ('load_from_id', 'bandit_scale/0'), ('load_from_id', 'memory_len/5'),
    ('load_from_id', 'memory_size/16'), ('load_from_id', 'umbrella_distract/0'),
    ('class', ('umbrella_chain', dict(chain_length=12, n_distractor=20, seed=0))),
    ('class', ('catch', dict(seed=0))),
    ('class', ('deep_sea', dict(size=10, deterministic=False, seed=0, mapping_seed=42))),
    ('class', ('cartpole', dict(seed=0))),
    ('class', ('cartpole_swingup', dict(seed=0))),
    ('class', ('mountain_car', dict(seed=0))),
]


def num_actions_of(env) -> int:
  return int(env.action_spec().num_values)


def build_case(c, mnist_dir):
  lanes = list(range(NUM_LANES))
  envs = [rr.make_reference_env(c['env_class'], c['kwargs'], c['rng'], c['seed'], lane, c['wrapper'],
                                c['wrapper_arg'], mnist_dir) for lane in lanes]
  n_actions = num_actions_of(envs[0])
  # crc32 is stable across runs (hash() is not)
  import zlib
  action_rng = np.random.RandomState(zlib.crc32(c['name'].encode()) % (2**31))
  actions = action_rng.randint(n_actions, size=(c['steps'], NUM_LANES)).astype(np.int32)
  reset_at = set(c['reset_at'])
  step_type = np.zeros((c['steps'], NUM_LANES), np.int32)
  reward = np.zeros((c['steps'], NUM_LANES), np.float64)
  discount = np.zeros((c['steps'], NUM_LANES), np.float64)
  obs = None
  for i, env in enumerate(envs):
    for t in range(c['steps']):
      ts = env.reset() if t in reset_at else env.step(int(actions[t, i]))
      o = np.asarray(ts.observation)
      assert o.dtype == np.float32, o.dtype
      if obs is None:
        obs = np.zeros((c['steps'], NUM_LANES) + o.shape, np.float32)
      step_type[t, i] = int(ts.step_type)
      reward[t, i] = np.nan if ts.reward is None else float(ts.reward)
      discount[t, i] = np.nan if ts.discount is None else float(ts.discount)
      obs[t, i] = o
  info_names = sorted(envs[0].bsuite_info().keys())
  info = np.array([[float(env.bsuite_info()[k]) for k in info_names] for env in envs], np.float64).reshape(
      NUM_LANES, len(info_names))
  meta = dict(c, num_actions=n_actions, info_names=info_names, lanes=lanes,
              numpy_version=np.__version__, mnist=dict(seed=MNIST_SEED, num_train=MNIST_TRAIN, num_test=MNIST_TEST))
  return dict(meta=np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8), actions=actions,
              step_type=step_type, reward=reward, discount=discount, observation=obs, info=info)


def build_known_answers(mnist_dir):
  del mnist_dir
  bsuite = rr.import_reference()
  rows = []
  for kind, what in KNOWN_ANSWERS:
    if kind == 'load_from_id':
      env = bsuite.load_from_id(what)
      label = what
      spec = dict(kind=kind, bsuite_id=what)
    else:
      env_class, kwargs = what
      import importlib
      module_name, class_name = rr._CLASSES[env_class]  # pylint: disable=protected-access
      env = getattr(importlib.import_module(module_name), class_name)(**kwargs)
      label = f'{env_class}({kwargs})'
      spec = dict(kind=kind, env_class=env_class, kwargs=kwargs)
    n_actions = num_actions_of(env)
    actions = np.random.RandomState(0).randint(n_actions, size=1000)
    trace = rr.run_trace(env, actions, explicit_reset=True)
    rows.append(dict(spec, label=label, digest=rr.trace_digest(trace),
                     reward_sum=float(np.nansum(trace['reward'])), num_last=int((trace['step_type'] == 2).sum()),
                     info=trace['info'],
                     reward_f32_sum=float(np.nansum(trace['reward'].astype(np.float32).astype(np.float64)))))
  return rows


def main():
  parser = argparse.ArgumentParser()
  parser.add_argument('--check', action='store_true')
  parser.add_argument('--only', default=None)
  args = parser.parse_args()
  os.makedirs(GOLDEN_DIR, exist_ok=True)
  from bsuite_b200 import datasets  # writer of the synthetic idx files (format only; no dynamics)
  mnist_dir = tempfile.mkdtemp(prefix='bsb_mnist_')
  datasets.write_synthetic_mnist(mnist_dir, MNIST_TRAIN, MNIST_TEST, MNIST_SEED)
  failures = 0
  for c in CASES:
    if args.only and args.only not in c['name']:
      continue
    data = build_case(c, mnist_dir)
    path = os.path.join(GOLDEN_DIR, c['name'] + '.npz')
    if args.check:
      old = np.load(path)
      same = all(np.array_equal(old[k], data[k], equal_nan=True) if k != 'meta' else True for k in data)
      print(('ok   ' if same else 'DIFF ') + c['name'])
      failures += not same
    else:
      np.savez_compressed(path, **data)
      print(f'wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)')
  if not args.only:
    rows = build_known_answers(mnist_dir)
    path = os.path.join(GOLDEN_DIR, 'known_answers.json')
    if args.check:
      old = json.load(open(path))
      same = old == json.loads(json.dumps(rows))
      print(('ok   ' if same else 'DIFF ') + 'known_answers.json')
      failures += not same
    else:
      json.dump(rows, open(path, 'w'), indent=1, sort_keys=True)
      print(f'wrote {path}')
  return failures


if __name__ == '__main__':
  sys.exit(main())
