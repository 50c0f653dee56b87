"""Batched counterparts of the reference's agent loop and trajectory buffer.

  * `run(agent, environment, num_steps)` -- `bsuite/baselines/experiment.py:24-57` for B lanes in lock-step: the
    agent sees the whole batch (`select_action(timestep) -> int tensor [B]`, `update(timestep, actions,
    new_timestep)`); lanes reset themselves, so the loop is over steps, not episodes.  The reference loop itself
    runs unchanged on the B = 1 face (`DmEnvAdapter`): it only calls `reset()` / `step()`.
  * `RandomAgent` -- `bsuite/baselines/random/agent.py:26-45` with one generator call per step for the whole batch.
  * `Trajectory` / `collect` -- the `[T + 1]` observations / `[T]` actions, rewards, discounts layout of
    `bsuite/baselines/utils/sequence.py:26-35`, as device tensors with a lane axis, filled by ONE fused rollout
    (`bsb_rollout`: on-device uniform random actions) instead of T appends.
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
