"""CPU restatement of the reference's per-environment step()/reset() dynamics.

TEST INFRASTRUCTURE ONLY.  This module is the checker the CUDA path is compared
against where the reference itself cannot run (the GPU box has no
/root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; nothing under bsuite_b200/
does.

Pinning: tests/test_oracle_pinned.py checks every function here, for every
fixture under tests/golden/, against traces recorded from the UNMODIFIED
reference by oracle/gen_golden.py (which imports /root/reference directly), and
against the known-answer digests of SURVEY.md 8c (tests/golden/known_answers.json).

The algorithm lives partly in a third-party dependency of the reference: numpy
(setup.py:85, unpinned; 2.3.5 here) supplies `RandomState` and the scalar
cos/sin/remainder/clip.  The oracle calls numpy for exactly those pieces, like
the reference does, so they are not restated.

Shape of the restatement: one `OracleEnv` object = one environment instance.
State is a plain dict; each family contributes three functions
(`_begin_*` = episode start, `_advance_*` = one transition, `_render_*` = the
observation), and `OracleEnv.step/reset` implement the auto-reset gate of
`bsuite/environments/base.py:54-65` plus the reward wrappers of
`bsuite/utils/wrappers.py:272-283,335-346`.
"""

from typing import Any, Dict, Optional, Tuple

import numpy as np

FIRST, MID, LAST = 0, 1, 2


def _mid(reward, obs):
  return MID, reward, 1.0, obs           # dm_env.transition


def _last(reward, obs):
  return LAST, reward, 0.0, obs          # dm_env.termination


# --------------------------------------------------------------------------- deep_sea
def _setup_deep_sea(s, size, deterministic=True, unscaled_move_cost=0.01, randomize_actions=True,
                    mapping_seed=None):
  """deep_sea.py:51-101."""
  s.update(n=size, det=deterministic, cost=unscaled_move_cost)
  if randomize_actions:
    s['mapping'] = np.random.RandomState(mapping_seed).binomial(1, 0.5, [size, size])   # :80-81
  else:
    s['mapping'] = np.ones([size, size])                                               # :85
  s.update(row=0, col=0, bad=False, total_bad_episodes=0, denoised_return=0)
  s['obs_shape'] = (size, size)
  s['num_actions'] = 2


def _render_deep_sea(s, rng):
  """deep_sea.py:103-108: one-hot of (row, col); all zeros once row == N."""
  grid = np.zeros((s['n'], s['n']), np.float32)
  if s['row'] < s['n']:
    grid[s['row'], s['col']] = 1.
  return grid


def _begin_deep_sea(s, rng):
  s.update(row=0, col=0, bad=False)                                                   # :110-114


def _advance_deep_sea(s, action, rng):
  """deep_sea.py:116-144, in the reference's evaluation order."""
  n, row, col = s['n'], s['row'], s['col']
  went_right = action == s['mapping'][row, col]
  r = 0.
  if col == n - 1 and went_right:
    r += 1.
    s['denoised_return'] += 1.
  if not s['det'] and row == n - 1 and col in (0, n - 1):
    r += rng.randn()
  if went_right:
    # rand() is evaluated before `or deterministic` (:130)
    if rng.rand() > 1 / n or s['det']:
      s['col'] = min(col + 1, n - 1)
    r -= s['cost'] / n
  else:
    if row == col:
      s['bad'] = True
    s['col'] = max(col - 1, 0)
  s['row'] = row + 1
  if s['row'] == n:
    s['total_bad_episodes'] += int(s['bad'])
    return LAST, r
  return MID, r


# --------------------------------------------------------------------------- catch
def _setup_catch(s, rows=10, columns=5):
  s.update(rows=rows, cols=columns, ball_x=None, ball_y=None, paddle_x=None, total_regret=0.)   # catch.py:45-66
  s['obs_shape'] = (rows, columns)
  s['num_actions'] = 3


def _render_catch(s, rng):
  board = np.zeros((s['rows'], s['cols']), np.float32)                                # catch.py:109-114
  board[s['ball_y'], s['ball_x']] = 1.
  board[s['rows'] - 1, s['paddle_x']] = 1.
  return board


def _begin_catch(s, rng):
  s['ball_x'] = rng.randint(s['cols'])                                                # catch.py:71
  s['ball_y'] = 0
  s['paddle_x'] = s['cols'] // 2


def _advance_catch(s, action, rng):
  s['paddle_x'] = int(np.clip(s['paddle_x'] + (action - 1), 0, s['cols'] - 1))        # :84-85
  s['ball_y'] += 1
  if s['ball_y'] == s['rows'] - 1:                                                    # :91-95
    r = 1. if s['paddle_x'] == s['ball_x'] else -1.
    s['total_regret'] += 1. - r
    return LAST, r
  return MID, 0.


# --------------------------------------------------------------------------- cartpole (+ swingup)
_POLE = dict(mass_cart=1., mass_pole=0.1, length=0.5, force_mag=10., gravity=9.8)      # cartpole.py:106-112


def _pole_physics(s, action):
  """cartpole.py:37-65: explicit Euler step from the old state."""
  x, x_dot, th, th_dot, t = s['x'], s['x_dot'], s['theta'], s['theta_dot'], s['t']
  dt = s['dt']
  force = (action - 1) * _POLE['force_mag']
  c, sn = np.cos(th), np.sin(th)
  pl = _POLE['mass_pole'] * _POLE['length']
  m_total = _POLE['mass_cart'] + _POLE['mass_pole']
  temp = (force + pl * th_dot**2 * sn) / m_total
  th_acc = (_POLE['gravity'] * sn - c * temp) / (_POLE['length'] * (4 / 3 - _POLE['mass_pole'] * c**2 / m_total))
  x_acc = temp - pl * th_acc * c / m_total
  s['x'] = x + dt * x_dot
  s['x_dot'] = x_dot + dt * x_acc
  s['theta'] = np.remainder(th + dt * th_dot, 2 * np.pi)
  s['theta_dot'] = th_dot + dt * th_acc
  s['t'] = t + dt


def _setup_cartpole(s, height_threshold=0.8, x_threshold=3., timescale=0.01, max_time=10., init_range=0.05):
  s.update(h=height_threshold, x_thr=x_threshold, dt=timescale, t_max=max_time, init=init_range,
           x=0, x_dot=0, theta=0, theta_dot=0, t=0, raw_return=0., best_episode=0., episode_return=0.)
  s['obs_shape'] = (1, 6)
  s['num_actions'] = 3


def _draw_pole_start(s, rng, offset):
  u = lambda: rng.uniform(low=-s['init'], high=s['init'])                             # cartpole.py:92
  s['x'], s['x_dot'] = u(), u()
  s['theta'] = offset + u()
  s['theta_dot'] = u()
  s['t'] = 0.


def _begin_cartpole(s, rng):
  _draw_pole_start(s, rng, 0)                                                         # cartpole.py:118-128
  s['episode_return'] = 0


def _pole_obs(s, width):
  o = np.zeros((1, width), np.float32)                                                # cartpole.py:167-177
  o[0, 0] = s['x'] / s['x_thr']
  o[0, 1] = s['x_dot'] / s['x_thr']
  o[0, 2] = np.sin(s['theta'])
  o[0, 3] = np.cos(s['theta'])
  o[0, 4] = s['theta_dot']
  o[0, 5] = s['t'] / s['t_max']
  return o


def _render_cartpole(s, rng):
  return _pole_obs(s, 6)


def _advance_cartpole(s, action, rng):
  _pole_physics(s, action)
  ok = np.cos(s['theta']) > s['h'] and np.abs(s['x']) < s['x_thr']                    # cartpole.py:140-153
  r = 1. if ok else 0.
  s['raw_return'] += r
  s['episode_return'] += r
  if s['t'] > s['t_max'] or not ok:
    s['best_episode'] = max(s['episode_return'], s['best_episode'])
    return LAST, r
  return MID, r


def _setup_cartpole_swingup(s, height_threshold=0.5, theta_dot_threshold=1., x_reward_threshold=1., move_cost=0.1,
                            x_threshold=3., timescale=0.01, max_time=10., init_range=0.05):
  _setup_cartpole(s, height_threshold, x_threshold, timescale, max_time, init_range)
  s.update(thd_thr=theta_dot_threshold, x_rew=x_reward_threshold, move_cost=move_cost, total_upright=0.)
  s['obs_shape'] = (1, 8)


def _begin_cartpole_swingup(s, rng):
  _draw_pole_start(s, rng, np.pi)                                                     # cartpole_swingup.py:81-91
  s['episode_return'] = 0.


def _render_cartpole_swingup(s, rng):
  o = _pole_obs(s, 8)                                                                 # cartpole_swingup.py:137-150
  o[0, 6] = 1. if np.abs(s['x']) < s['x_rew'] else -1.
  o[0, 7] = 1. if np.abs(s['theta_dot']) < s['thd_thr'] else -1.
  return o


def _advance_cartpole_swingup(s, action, rng):
  _pole_physics(s, action)
  upright = (np.cos(s['theta']) > s['h'] and np.abs(s['theta_dot']) < s['thd_thr']
             and np.abs(s['x']) < s['x_rew'])                                         # cartpole_swingup.py:104-107
  r = -1. * np.abs(action - 1) * s['move_cost']
  if upright:
    r += 1.
    s['total_upright'] += 1
  s['raw_return'] += r
  s['episode_return'] += r
  if s['t'] > s['t_max'] or np.abs(s['x']) > s['x_thr']:                              # :116-121
    s['best_episode'] = max(s['episode_return'], s['best_episode'])
    return LAST, r
  return MID, r


# --------------------------------------------------------------------------- mountain_car
def _setup_mountain_car(s, max_steps=1000):
  s.update(max_steps=max_steps, tick=0, raw_return=0., pos=0., vel=0.)                # mountain_car.py:36-60
  s['obs_shape'] = (1, 3)
  s['num_actions'] = 3


def _render_mountain_car(s, rng):
  return np.array([[s['pos'], s['vel'], s['tick'] / s['max_steps']]], dtype=np.float32)   # :62-64


def _begin_mountain_car(s, rng):
  s['tick'] = 0                                                                       # :66-71
  s['pos'] = rng.uniform(-0.6, -0.4)
  s['vel'] = 0


def _advance_mountain_car(s, action, rng):
  s['tick'] += 1                                                                      # :73-90
  s['raw_return'] += -1.
  s['vel'] += (action - 1) * 0.001 + np.cos(3 * s['pos']) * -0.0025
  s['vel'] = np.clip(s['vel'], -0.07, 0.07)
  s['pos'] += s['vel']
  s['pos'] = np.clip(s['pos'], -1.2, 0.6)
  if s['pos'] == -1.2:
    s['vel'] = np.clip(s['vel'], 0, 0.07)
  done = s['pos'] >= 0.5 or s['tick'] >= s['max_steps']
  return (LAST if done else MID), -1.


# --------------------------------------------------------------------------- memory_chain
def _setup_memory_chain(s, memory_length, num_bits=1):
  s.update(length=memory_length, bits=num_bits, tick=0, total_perfect=0, total_regret=0)
  s['obs_shape'] = (1, num_bits + 2)
  s['num_actions'] = 2
  s['ctor_draws'] = True          # memory_chain.py:49-50 draws a context/query nobody sees


def _draw_memory(s, rng):
  s['context'] = rng.binomial(1, 0.5, s['bits'])
  s['query'] = rng.randint(s['bits'])


def _render_memory_chain(s, rng):
  o = np.zeros((1, s['bits'] + 2), np.float32)                                        # memory_chain.py:60-71
  o[0, 0] = 1 - s['tick'] / s['length']
  if s['tick'] == s['length'] - 1:
    o[0, 1] = s['query']
  if s['tick'] == 0:
    o[0, 2:] = 2 * s['context'] - 1
  return o


def _begin_memory_chain(s, rng):
  s['tick'] = 0                                                                       # :91-97
  _draw_memory(s, rng)


def _advance_memory_chain(s, action, rng):
  # NB the observation of this transition is rendered BEFORE the tick (:74-75); OracleEnv handles that
  # through `render_before_advance`.
  s['tick'] += 1
  if s['tick'] - 1 < s['length']:
    return MID, 0.
  if action == s['context'][s['query']]:                                              # :83-88
    s['total_perfect'] += 1
    return LAST, 1.
  s['total_regret'] += 2.
  return LAST, -1.


# --------------------------------------------------------------------------- bandit
def _setup_bandit(s, mapping_seed=None, num_actions=11):
  rng = np.random.RandomState(mapping_seed)                                           # bandit.py:43-47
  order = rng.choice(range(num_actions), size=num_actions, replace=False)
  s.update(rewards=np.linspace(0, 1, num_actions)[order], total_regret=0.)
  s['obs_shape'] = (1, 1)
  s['num_actions'] = num_actions


def _render_bandit(s, rng):
  return np.ones((1, 1), np.float32)                                                  # bandit.py:53-54


def _begin_bandit(s, rng):
  pass


def _advance_bandit(s, action, rng):
  r = s['rewards'][action]                                                            # bandit.py:60-64
  s['total_regret'] += 1. - r
  return LAST, r


# --------------------------------------------------------------------------- umbrella_chain
def _setup_umbrella_chain(s, chain_length, n_distractor=0):
  s.update(length=chain_length, n=n_distractor, tick=0, has=0, total_regret=0)
  s['obs_shape'] = (1, 3 + n_distractor)
  s['num_actions'] = 2
  s['ctor_draws'] = True          # umbrella_chain.py:55 draws need_umbrella once at construction


def _render_umbrella_chain(s, rng):
  o = np.zeros((1, 3 + s['n']), np.float32)                                           # umbrella_chain.py:60-66
  o[0, 0] = s['need']
  o[0, 1] = s['has']
  o[0, 2] = 1 - s['tick'] / s['length']
  o[0, 3:] = rng.binomial(1, 0.5, size=s['n'])       # fresh distractors on every call
  return o


def _begin_umbrella_chain(s, rng):
  s['tick'] = 0                                                                       # :87-92
  s['need'] = rng.binomial(1, 0.5)
  s['has'] = rng.binomial(1, 0.5)


def _advance_umbrella_chain(s, action, rng):
  s['tick'] += 1                                                                      # :68-85
  if s['tick'] == 1:
    s['has'] = action
  if s['tick'] == s['length']:
    if s['has'] == s['need']:
      return LAST, 1.
    s['total_regret'] += 2.
    return LAST, -1.
  return MID, 2. * rng.binomial(1, 0.5) - 1.        # drawn before the observation's distractors


# --------------------------------------------------------------------------- discounting_chain
_REWARD_TICKS = (1, 3, 10, 30, 100)                                                    # discounting_chain.py:47


def _setup_discounting_chain(s, mapping_seed=None):
  if mapping_seed is None:
    mapping_seed = np.random.randint(0, 5)
  rewards = np.ones(5)
  rewards[mapping_seed % 5] += 0.1                                                    # :52-56
  s.update(rewards=rewards, tick=0, context=-1)
  s['obs_shape'] = (1, 2)
  s['num_actions'] = 5


def _render_discounting_chain(s, rng):
  o = np.zeros((1, 2), np.float32)                                                    # :63-67
  o[0, 0] = s['context']
  o[0, 1] = s['tick'] / 100
  return o


def _begin_discounting_chain(s, rng):
  s.update(tick=0, context=-1)


def _advance_discounting_chain(s, action, rng):
  if s['tick'] == 0:
    s['context'] = action                                                             # :76-77
  s['tick'] += 1
  r = s['rewards'][s['context']] if s['tick'] == _REWARD_TICKS[s['context']] else 0.
  return (LAST if s['tick'] == 100 else MID), r


# --------------------------------------------------------------------------- mnist
def _setup_mnist(s, fraction=1., images=None, labels=None):
  count = int(fraction * len(labels))                                                 # mnist.py:46-52
  s.update(images=images[:count], labels=labels[:count], count=count, total_regret=0., shown=None)
  s['obs_shape'] = tuple(images.shape[1:])
  s['num_actions'] = 10


def _render_mnist(s, rng):
  if s['shown'] is None:
    return np.zeros(s['obs_shape'], np.float32)                                       # mnist.py:74
  return s['images'][s['shown']].astype(np.float32) / 255                             # mnist.py:64 (int8 pixels!)


def _begin_mnist(s, rng):
  s['shown'] = rng.randint(s['count'])                                                # mnist.py:63
  s['label'] = s['labels'][s['shown']]


def _advance_mnist(s, action, rng):
  r = 1. if action == s['label'] else -1.                                             # mnist.py:71-73
  s['total_regret'] += 1. - r
  s['shown'] = None
  return LAST, r


_FAMILIES = {
    'deep_sea': (_setup_deep_sea, _begin_deep_sea, _advance_deep_sea, _render_deep_sea,
                 ('total_bad_episodes', 'denoised_return')),
    'catch': (_setup_catch, _begin_catch, _advance_catch, _render_catch, ('total_regret',)),
    'cartpole': (_setup_cartpole, _begin_cartpole, _advance_cartpole, _render_cartpole, ('raw_return', 'best_episode')),
    'cartpole_swingup': (_setup_cartpole_swingup, _begin_cartpole_swingup, _advance_cartpole_swingup,
                         _render_cartpole_swingup, ('raw_return', 'total_upright', 'best_episode')),
    'mountain_car': (_setup_mountain_car, _begin_mountain_car, _advance_mountain_car, _render_mountain_car,
                     ('raw_return',)),
    'memory_chain': (_setup_memory_chain, _begin_memory_chain, _advance_memory_chain, _render_memory_chain,
                     ('total_perfect', 'total_regret')),
    'bandit': (_setup_bandit, _begin_bandit, _advance_bandit, _render_bandit, ('total_regret',)),
    'umbrella_chain': (_setup_umbrella_chain, _begin_umbrella_chain, _advance_umbrella_chain,
                       _render_umbrella_chain, ('total_regret',)),
    'discounting_chain': (_setup_discounting_chain, _begin_discounting_chain, _advance_discounting_chain,
                          _render_discounting_chain, ()),
    'mnist': (_setup_mnist, _begin_mnist, _advance_mnist, _render_mnist, ('total_regret',)),
}


def philox_random_state(seed: int, lane: int, stream: int = 0) -> np.random.RandomState:
  """numpy's legacy RandomState over the Philox stream engine lane `lane` consumes."""
  bitgen = np.random.Philox(key=np.array([seed, lane], dtype=np.uint64),
                            counter=np.array([0, 0, 0, stream], dtype=np.uint64))
  return np.random.RandomState(bitgen)


class OracleEnv:
  """One environment instance; mirrors base.Environment.step/reset + the reward wrappers."""

  def __init__(self, env_class: str, kwargs: Optional[Dict[str, Any]] = None, rng: str = 'philox', seed: int = 0,
               lane: int = 0, wrapper: Optional[str] = None, wrapper_arg: float = 0.0):
    setup, self._begin, self._advance, self._render, self._info_names = _FAMILIES[env_class]
    self._render_before_advance = env_class == 'memory_chain'
    if rng == 'philox':
      self._rng = philox_random_state(seed, lane, 0)
      self._wrapper_rng = philox_random_state(seed, lane, 1)
    elif rng == 'mt19937':
      self._rng = np.random.RandomState((seed + lane) % 2**32)
      self._wrapper_rng = np.random.RandomState((seed + lane) % 2**32)                # wrappers.py:267: same seed
    else:
      raise ValueError(rng)
    self._wrapper, self._wrapper_arg = wrapper, wrapper_arg
    self.state: Dict[str, Any] = {}
    setup(self.state, **(kwargs or {}))
    if self.state.get('ctor_draws'):
      if env_class == 'memory_chain':
        _draw_memory(self.state, self._rng)
      else:
        self.state['need'] = self._rng.binomial(1, 0.5)
    self._needs_reset = True                                                          # base.py:51-52
    self.obs_shape = self.state['obs_shape']
    self.num_actions = self.state['num_actions']

  def reset(self) -> Tuple[int, Optional[float], Optional[float], np.ndarray]:
    self._needs_reset = False                                                         # base.py:54-57
    self._begin(self.state, self._rng)
    return FIRST, None, None, self._render(self.state, self._rng)

  def step(self, action: int):
    if self._needs_reset:                                                             # base.py:61-62
      return self.reset()
    if self._render_before_advance:
      obs = self._render(self.state, self._rng)
      step_type, reward = self._advance(self.state, action, self._rng)
    else:
      step_type, reward = self._advance(self.state, action, self._rng)
      obs = self._render(self.state, self._rng)
    self._needs_reset = step_type == LAST                                             # base.py:64
    if self._wrapper == 'noise':                                                      # wrappers.py:275-283
      reward = reward + self._wrapper_arg * self._wrapper_rng.randn()
    elif self._wrapper == 'scale':                                                    # wrappers.py:338-346
      reward = reward * self._wrapper_arg
    return step_type, reward, (0.0 if step_type == LAST else 1.0), obs

  def bsuite_info(self) -> Dict[str, float]:
    return {k: self.state[k] for k in self._info_names}


def run_lanes(env_class, kwargs, actions: np.ndarray, rng='philox', seed=0, lane_offset=0, wrapper=None,
              wrapper_arg=0.0, reset_at=()) -> Dict[str, np.ndarray]:
  """Runs `actions` [T, L] through L oracle lanes; FIRST rows carry reward = discount = 0 like the engine."""
  T, L = actions.shape
  envs = [OracleEnv(env_class, kwargs, rng, seed, lane_offset + i, wrapper, wrapper_arg) for i in range(L)]
  step_type = np.zeros((T, L), np.int32)
  reward = np.zeros((T, L), np.float64)
  discount = np.zeros((T, L), np.float32)
  obs = np.zeros((T, L) + tuple(envs[0].obs_shape), np.float32)
  reset_at = set(reset_at)
  for i, env in enumerate(envs):
    for t in range(T):
      st, r, d, o = env.reset() if t in reset_at else env.step(int(actions[t, i]))
      step_type[t, i] = st
      reward[t, i] = 0.0 if r is None else r
      discount[t, i] = 0.0 if d is None else d
      obs[t, i] = o
  names = envs[0]._info_names  # pylint: disable=protected-access
  info = {k: np.array([float(e.bsuite_info()[k]) for e in envs]) for k in names}
  return dict(step_type=step_type, reward=reward, discount=discount, observation=obs, info=info)
