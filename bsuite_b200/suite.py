"""Heterogeneous batches: many bsuite_ids at once (BASELINE config #5).

`SweepBatch` holds one `BatchedEnvironment` per bsuite_id, each with `lanes` lanes, and advances all of them
"in lock-step" from the caller's point of view: every environment's fused rollout is enqueued on its own CUDA
stream, so the 23 small kernels of a full-sweep step overlap on the GPU instead of queueing behind each other.
Across GPUs every id's lanes are sharded evenly (rank r owns lanes [r*lanes/W, (r+1)*lanes/W) of EVERY id), so the
observation-heavy families (deep_sea, mnist) do not imbalance the ranks; the only collective is the all-gather of
per-rank return statistics at log points.
"""

from typing import Dict, List, Optional, Sequence

from bsuite_b200 import distributed
from bsuite_b200 import registry
from bsuite_b200 import sweep


def one_per_experiment(setting: int = 0) -> List[str]:
  """`<experiment>/<setting>` for each of the 23 experiments (the sweep.TESTING idea, including noise/scale)."""
  return [ids[min(setting, len(ids) - 1)] for ids in sweep.BY_EXPERIMENT.values()]


class SweepBatch:

  def __init__(self, bsuite_ids: Optional[Sequence[str]] = None, lanes: int = 4096, device='cuda', seed: int = 0,
               rank: int = 0, world: int = 1, track_episodes: bool = True):
    import torch
    self._torch = torch
    self.bsuite_ids = list(bsuite_ids) if bsuite_ids is not None else one_per_experiment()
    first, count = distributed.shard_range(lanes, rank, world)
    self.lanes, self.local_lanes, self.lane_offset = lanes, count, first
    self.envs = {
        bsuite_id: registry.load_from_id(bsuite_id, batch=count, device=device, seed=seed, lane_offset=first,
                                         track_episodes=track_episodes)
        for bsuite_id in self.bsuite_ids
    }
    self._device = next(iter(self.envs.values())).device
    self._cuda = self._device.type == 'cuda'
    self._streams = {k: torch.cuda.Stream(device=self._device) for k in self.envs} if self._cuda else {}
    self._buffers: Dict[str, object] = {}
    self._buffer_steps = None

  def _ensure_buffers(self, num_steps: int):
    if self._buffer_steps != num_steps:
      self._buffers = {k: env.make_buffers(num_steps, with_actions=True) for k, env in self.envs.items()}
      self._buffer_steps = num_steps

  def rollout(self, num_steps: int, action_seed: int = 0):
    """`num_steps` fused steps of every environment (on-device uniform random actions); returns id -> TimeStep.

    The returned tensors are reused by the next call.  On CUDA each environment runs on its own stream; the
    caller's current stream waits for all of them before this function returns control of the outputs.
    """
    torch = self._torch
    self._ensure_buffers(num_steps)
    result = {}
    if not self._cuda:
      for k, env in self.envs.items():
        result[k] = env.rollout(num_steps, action_seed=action_seed, out=self._buffers[k])
      return result
    current = torch.cuda.current_stream(self._device)
    for k, env in self.envs.items():
      stream = self._streams[k]
      stream.wait_stream(current)
      with torch.cuda.stream(stream):
        result[k] = env.rollout(num_steps, action_seed=action_seed, out=self._buffers[k])
    for stream in self._streams.values():
      current.wait_stream(stream)
    return result

  def local_returns(self):
    """float64 [n_ids, 3] on the device: per-id sums of (total_return, episode, steps) over this rank's lanes."""
    torch = self._torch
    sums = torch.stack([env.episode_stat_sums() for env in self.envs.values()])   # one reduction kernel per id
    return sums[:, [2, 1, 0]]

  def gather_returns(self):
    """The one collective of the path: all-gather of `local_returns()`; returns [world, n_ids, 3]."""
    torch = self._torch
    import torch.distributed as dist
    block = self.local_returns()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
      world = dist.get_world_size()
      flat = block.contiguous().reshape(-1)        # 1-D in, 1-D out: the one layout every backend accepts
      out = torch.empty(world * flat.numel(), dtype=block.dtype, device=block.device)
      dist.all_gather_into_tensor(out, flat)
      return out.reshape((world,) + tuple(block.shape))
    return block.unsqueeze(0)

  def bytes_per_step(self) -> int:
    """Algorithmic bytes of one lock-step of the whole local batch (obs + 16 B of scalars per lane; state excluded)."""
    total = 0
    for env in self.envs.values():
      numel = 1
      for d in env.obs_shape:
        numel *= d
      total += env.batch * (4 * numel + 16)
    return total

  def close(self):
    for env in self.envs.values():
      env.close()
