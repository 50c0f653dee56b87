"""Batched counterparts of the reference's agent loop and trajectory buffer.

  * `run(agent, environment, num_steps)` -- `bsuite/baselines/experiment.py:24-57` for B lanes in lock-step: the
    agent sees the whole batch (`select_action(timestep) -> int tensor [B]`, `update(timestep, actions,
    new_timestep)`); lanes reset themselves, so the loop is over steps, not episodes.  The reference loop itself
    runs unchanged on the B = 1 face (`DmEnvAdapter`): it only calls `reset()` / `step()`.
  * `RandomAgent` -- `bsuite/baselines/random/agent.py:26-45` with one generator call per step for the whole batch.
  * `Trajectory` / `collect` -- the `[T + 1]` observations / `[T]` actions, rewards, discounts layout of
    `bsuite/baselines/utils/sequence.py:26-35`, as device tensors with a lane axis, filled by ONE fused rollout
    (`bsb_rollout`: on-device uniform random actions) instead of T appends.
  * `HostHalves` -- the reference's strict host loop (one decision per `env.step`, experiment.py:45-57) over two
    half-batches driven alternately, so that one half's PCIe round trip and decision hide behind the other half's
    kernel.
"""

from typing import Any, NamedTuple, Optional


class Trajectory(NamedTuple):
  """T transitions of B lanes.  `observations[t]` is what the agent saw before `actions[t]`;
  `rewards[t]`, `discounts[t]`, `step_types[t]` belong to the timestep that followed.  A lane whose `step_types[t]`
  is FIRST (0) restarted at that call: its action was ignored and reward / discount are 0 (the reference: None)."""
  observations: Any   # [T + 1, B, ...] float32
  actions: Any        # [T, B] int32
  rewards: Any        # [T, B]
  discounts: Any      # [T, B] float32
  step_types: Any     # [T, B] int32


def collect(environment, num_steps: int, action_seed: int = 0, last_observation=None) -> Trajectory:
  """One fused rollout of `num_steps` steps with on-device random actions, returned in Trajectory layout.

  `last_observation` [B, ...] is the observation the lanes showed before this call (the previous trajectory's
  `observations[-1]`); on a fresh environment it is not needed: the first call of every lane returns FIRST.
  """
  import torch
  out = environment.make_buffers(num_steps, with_actions=True)
  ts = environment.rollout(num_steps, action_seed=action_seed, out=out)
  if last_observation is None:
    last_observation = torch.zeros_like(ts.observation[0])
  observations = torch.cat([last_observation.unsqueeze(0), ts.observation], dim=0)
  return Trajectory(observations, out.actions, ts.reward, ts.discount, ts.step_type)


class RandomAgent:
  """Uniform random actions for every lane (baselines/random/agent.py:26-45)."""

  def __init__(self, action_spec, batch: int, device='cuda', seed: Optional[int] = None):
    import torch
    self._torch = torch
    self._num_actions, self._batch = int(action_spec.num_values), int(batch)
    self._device = torch.device(device)
    self._generator = torch.Generator(device=self._device)
    if seed is not None:
      self._generator.manual_seed(int(seed))

  def select_action(self, timestep):
    del timestep
    return self._torch.randint(0, self._num_actions, (self._batch,), dtype=self._torch.int32, device=self._device,
                               generator=self._generator)

  def update(self, timestep, action, new_timestep):
    del timestep, action, new_timestep


def run(agent, environment, num_steps: int) -> None:
  """Runs a batched agent on a batched environment for `num_steps` lock-steps (experiment.py:43-57)."""
  timestep = environment.reset()
  for _ in range(int(num_steps)):
    actions = agent.select_action(timestep)
    new_timestep = environment.step(actions)
    agent.update(timestep, actions, new_timestep)
    timestep = new_timestep


class Replay:
  """Uniform replay of flat item tuples as device tensors (`bsuite/baselines/utils/replay.py:24-88`): a ring of
  `capacity` slots per item, `add(items)` writes one tuple, `sample(size)` returns a list of `[size, ...]` tensors
  drawn uniformly with replacement, `size` / `fraction_filled` as in the reference.  `add_batch` writes many tuples
  at once (leading axis = tuple index) and `add_transitions` feeds it a whole `Trajectory`: the reference's DQN
  stores `(o_tm1, a_tm1, r_t, d_t, o_t)` per environment step (baselines/dqn/agent.py); here that is B x T tuples
  per fused rollout, minus the calls on which a lane merely restarted (step_type FIRST: no transition happened)."""

  def __init__(self, capacity: int, device='cuda', seed: Optional[int] = None):
    import torch
    self._torch = torch
    self._capacity, self._device = int(capacity), torch.device(device)
    self._data, self._num_added = None, 0
    self._generator = torch.Generator(device=self._device)
    if seed is not None:
      self._generator.manual_seed(int(seed))

  def _preallocate(self, items):
    torch = self._torch
    self._data = [torch.zeros((self._capacity,) + tuple(item.shape[1:]), dtype=item.dtype, device=self._device)
                  for item in items]

  def add(self, items) -> None:
    """Adds a single tuple of items (tensors or scalars, not batched)."""
    torch = self._torch
    self.add_batch([torch.as_tensor(item, device=self._device).unsqueeze(0) for item in items])

  def add_batch(self, items) -> None:
    """Adds `n` tuples at once: every item has a leading axis of length n; the oldest slots are overwritten."""
    torch = self._torch
    items = [torch.as_tensor(item, device=self._device) for item in items]
    n = int(items[0].shape[0])
    if n == 0:
      return
    if self._data is None:
      self._preallocate(items)
    if n > self._capacity:                         # only the newest `capacity` tuples can survive
      items = [item[n - self._capacity:] for item in items]
      self._num_added += n - self._capacity
      n = self._capacity
    slots = (torch.arange(n, device=self._device) + self._num_added) % self._capacity
    for slot, item in zip(self._data, items):
      slot[slots] = item.to(slot.dtype)
    self._num_added += n

  def add_transitions(self, trajectory: Trajectory) -> int:
    """Adds every real transition of a `Trajectory` as `(o_tm1, a_tm1, r_t, d_t, o_t)`; returns how many."""
    keep = (trajectory.step_types != 0).reshape(-1)
    flat = lambda x: x.reshape((-1,) + tuple(x.shape[2:]))[keep]
    o_tm1, o_t = trajectory.observations[:-1], trajectory.observations[1:]
    self.add_batch([flat(o_tm1), flat(trajectory.actions), flat(trajectory.rewards), flat(trajectory.discounts), flat(o_t)])
    return int(keep.sum())

  def sample(self, size: int):
    """A list of `[size, ...]` tensors, one per item, drawn uniformly with replacement from the filled slots."""
    indices = self._torch.randint(0, self.size, (int(size),), device=self._device, generator=self._generator)
    return [slot[indices] for slot in self._data]

  def reset(self) -> None:
    self._data, self._num_added = None, 0

  @property
  def size(self) -> int:
    return min(self._capacity, self._num_added)

  @property
  def fraction_filled(self) -> float:
    return self.size / self._capacity


class HostHalves:
  """`batch` lanes of one experiment as TWO half-batch environments that a HOST-side agent drives alternately.

  The reference's loop is strict: the agent acts on what the previous `env.step` returned (experiment.py:45-57).
  With the policy on the host, every step of one big batch leaves the GPU idle while reward / discount / step_type
  cross PCIe, the agent decides and the next actions come back.  Lanes are independent, so the batch
  is split in two handles (lane keys continue across the split: the trajectories are those of ONE `batch`-lane
  environment, bit for bit) and each half stays a strict loop of its own -- `submit(h, actions)` enqueues half h's
  step (`BSB_HOST_NO_WAIT`), `collect(h)` returns its host timestep -- while the OTHER half's kernel has the GPU:

      halves.submit(0, a0); halves.submit(1, a1)
      while ...:
        ts0, obs0 = halves.collect(0); halves.submit(0, policy(ts0))
        ts1, obs1 = halves.collect(1); halves.submit(1, policy(ts1))

  `run(policy, num_steps)` is that loop.  Needs a CUDA device (pinned buffers); host environments gain nothing
  from it and are refused.
  """

  def __init__(self, bsuite_id: str, batch: int, device='cuda', seed: Optional[int] = None, lane_offset: int = 0,
               **engine_kwargs):
    from bsuite_b200 import registry
    batch = int(batch)
    if batch < 2:
      raise ValueError('HostHalves needs at least two lanes')
    first = (batch // 2 + 31) // 32 * 32 if batch >= 64 else batch // 2     # whole warps in the first half
    kwargs = dict(engine_kwargs)
    if seed is not None:
      kwargs['seed'] = seed
    self.envs = [registry.load_from_id(bsuite_id, batch=first, device=device, lane_offset=lane_offset, **kwargs),
                 registry.load_from_id(bsuite_id, batch=batch - first, device=device, lane_offset=lane_offset + first,
                                       **kwargs)]
    if any(e.device.type != 'cuda' for e in self.envs):
      for e in self.envs:
        e.close()
      raise ValueError('HostHalves drives CUDA environments from pinned host buffers')
    self.batch, self.sizes = batch, (first, batch - first)
    self.host = [e.make_host_buffers() for e in self.envs]
    self.out = [e.make_buffers() for e in self.envs]
    self._in_flight = [False, False]

  def reset(self):
    """Resets both halves; returns their device TimeSteps."""
    self.drain()
    return [e.reset(out=o) for e, o in zip(self.envs, self.out)]

  def submit(self, half: int, actions):
    """Enqueues one step of `half` with `actions` (pinned CPU int32 [sizes[half]]); returns at once."""
    if self._in_flight[half]:
      raise RuntimeError(f'half {half} already has a step in flight: collect() it first')
    self.envs[half].step_host(actions, self.host[half], self.out[half], wait=False)
    self._in_flight[half] = True

  def collect(self, half: int):
    """Waits for the step of `half` in flight; returns (host TimeStep, device observation)."""
    self.envs[half].host_wait()
    self._in_flight[half] = False
    return self.host[half].timestep(), self.out[half].observation

  def drain(self):
    for half in (0, 1):
      if self._in_flight[half]:
        self.collect(half)

  def run(self, policy, num_steps: int, first_actions=None):
    """`num_steps` steps of every lane: `policy(half, step, host_timestep) -> pinned int32 actions` is asked once
    per half and step, always with that half's LATEST timestep (None before the first step unless
    `first_actions` = [actions0, actions1] is given).  Returns the final host timesteps of both halves."""
    last = [None, None]
    for half in (0, 1):
      self.submit(half, first_actions[half] if first_actions is not None else policy(half, 0, None))
    for step in range(1, int(num_steps)):
      for half in (0, 1):
        last[half] = self.collect(half)[0]
        self.submit(half, policy(half, step, last[half]))
    for half in (0, 1):
      last[half] = self.collect(half)[0]
    return last

  def close(self):
    self.drain()
    for e in self.envs:
      e.close()
