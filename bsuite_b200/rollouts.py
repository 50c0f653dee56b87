"""Batched counterparts of the reference's agent loop and trajectory buffer.

  * `run(agent, environment, num_steps)` -- `bsuite/baselines/experiment.py:24-57` for B lanes in lock-step: the
    agent sees the whole batch (`select_action(timestep) -> int tensor [B]`, `update(timestep, actions,
    new_timestep)`); lanes reset themselves, so the loop is over steps, not episodes.  The reference loop itself
    runs unchanged on the B = 1 face (`DmEnvAdapter`): it only calls `reset()` / `step()`.
  * `RandomAgent` -- `bsuite/baselines/random/agent.py:26-45` with one generator call per step for the whole batch.
  * `Trajectory` / `collect` -- the `[T + 1]` observations / `[T]` actions, rewards, discounts layout of
    `bsuite/baselines/utils/sequence.py:26-35`, as device tensors with a lane axis, filled by ONE fused rollout
    (`bsb_rollout`: on-device uniform random actions) instead of T appends.
  * `HostParts` / `HostHalves` -- the reference's strict host loop (one decision per `env.step`,
    experiment.py:45-57) over two (or more) part-batches driven round-robin, so that one part's PCIe round trip and
    decision hide behind the other parts' kernels.
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


class HostParts:
  """`batch` lanes of one experiment as `parts` environments that a HOST-side agent drives round-robin.

  The reference's loop is strict: the agent acts on what the previous `env.step` returned (experiment.py:45-57).
  With the policy on the host, every step of one big batch leaves the GPU idle while reward / discount / step_type
  cross PCIe, the agent decides and the next actions come back.  Lanes are independent, so the batch
  is split over `parts` handles (lane keys continue across the splits: the trajectories are those of ONE
  `batch`-lane environment, bit for bit) and each part stays a strict loop of its own -- `submit(p, actions)`
  enqueues part p's step (`BSB_HOST_NO_WAIT`), `collect(p)` returns its host timestep -- while the OTHER parts'
  kernels have the GPU:

      for p in range(parts): halves.submit(p, a[p])
      while ...:
        for p in range(parts):
          ts, obs = halves.collect(p); halves.submit(p, policy(ts))

  `run(policy, num_steps)` is that loop.  Two parts (`HostHalves`) hide one part's round trip behind the other's
  kernel; more parts give every round trip more kernels to hide behind at the price of more (smaller) launches and
  host calls per step.  Needs a CUDA device (pinned buffers); host environments gain nothing from it and are refused.
  """

  def __init__(self, bsuite_id: str, batch: int, device='cuda', seed: Optional[int] = None, lane_offset: int = 0,
               parts: int = 2, **engine_kwargs):
    from bsuite_b200 import registry
    batch, parts = int(batch), int(parts)
    if parts < 2:
      raise ValueError('HostParts needs at least two parts')
    if batch < parts:
      raise ValueError(f'HostParts needs at least one lane per part ({parts} parts, {batch} lanes)')
    self.sizes = split_sizes(batch, parts)
    kwargs = dict(engine_kwargs)
    if seed is not None:
      kwargs['seed'] = seed
    self.envs, offset = [], int(lane_offset)
    try:
      for size in self.sizes:
        self.envs.append(registry.load_from_id(bsuite_id, batch=size, device=device, lane_offset=offset, **kwargs))
        offset += size
      if any(e.device.type != 'cuda' for e in self.envs):
        raise ValueError('HostParts drives CUDA environments from pinned host buffers')
    except Exception:
      for e in self.envs:
        e.close()
      raise
    self.batch, self.parts = batch, parts
    self.host = [e.make_host_buffers() for e in self.envs]
    self.out = [e.make_buffers() for e in self.envs]
    self._in_flight = [False] * parts

  def reset(self):
    """Resets every part; returns their device TimeSteps."""
    self.drain()
    return [e.reset(out=o) for e, o in zip(self.envs, self.out)]

  def submit(self, part: int, actions):
    """Enqueues one step of `part` with `actions` (pinned CPU int32 [sizes[part]]); returns at once."""
    if self._in_flight[part]:
      raise RuntimeError(f'part {part} already has a step in flight: collect() it first')
    self.envs[part].step_host(actions, self.host[part], self.out[part], wait=False)
    self._in_flight[part] = True

  def collect(self, part: int):
    """Waits for the step of `part` in flight; returns (host TimeStep, device observation)."""
    self.envs[part].host_wait()
    self._in_flight[part] = False
    return self.host[part].timestep(), self.out[part].observation

  def drain(self):
    for part in range(self.parts):
      if self._in_flight[part]:
        self.collect(part)

  def run(self, policy, num_steps: int, first_actions=None):
    """`num_steps` steps of every lane: `policy(part, step, host_timestep) -> pinned int32 actions` is asked once
    per part and step, always with that part's LATEST timestep (None before the first step unless
    `first_actions` = [actions of part 0, ...] is given).  Returns the final host timesteps of the parts."""
    last = [None] * self.parts
    for part in range(self.parts):
      self.submit(part, first_actions[part] if first_actions is not None else policy(part, 0, None))
    for step in range(1, int(num_steps)):
      for part in range(self.parts):
        last[part] = self.collect(part)[0]
        self.submit(part, policy(part, step, last[part]))
    for part in range(self.parts):
      last[part] = self.collect(part)[0]
    return last

  def close(self):
    self.drain()
    for e in self.envs:
      e.close()


def split_sizes(batch: int, parts: int):
  """Lanes per part: whole warps (multiples of 32 lanes) in every part but the last when there are enough lanes,
  sizes as equal as that allows; `sum == batch`, every part non-empty."""
  batch, parts = int(batch), int(parts)
  if batch >= 64 * parts:
    warps = (batch + 31) // 32
    per, extra = divmod(warps, parts)
    sizes = [(per + (1 if i < extra else 0)) * 32 for i in range(parts)]
    sizes[-1] = batch - sum(sizes[:-1])
  else:
    per, extra = divmod(batch, parts)
    sizes = [per + (1 if i < extra else 0) for i in range(parts)]
  assert sum(sizes) == batch and all(size > 0 for size in sizes), (batch, parts, sizes)
  return tuple(sizes)


class HostHalves(HostParts):
  """`HostParts` with two parts: one half's PCIe round trip and decision hide behind the other half's kernel."""

  def __init__(self, bsuite_id: str, batch: int, device='cuda', seed: Optional[int] = None, lane_offset: int = 0,
               **engine_kwargs):
    if int(batch) < 2:
      raise ValueError('HostHalves needs at least two lanes')
    super().__init__(bsuite_id, batch, device=device, seed=seed, lane_offset=lane_offset, parts=2, **engine_kwargs)
