"""Multi-GPU: shard the lane axis, gather per-rank episode statistics.

Every lane is an independent MDP (one environment object per process in the reference,
baselines/utils/pool.py:48-51), so the batch axis shards with NO data-path collective.  Rank r of `world` owns
the contiguous global lanes [r * B/world, (r + 1) * B/world); RNG keys are functions of the GLOBAL lane id, so a
lane's trajectory does not depend on how the batch is sharded.  The only collective is one all-gather of a small
per-rank block of Logging statistics at log points (NCCL on GPUs; gloo in the CPU tests).
"""

from typing import Any, Dict, Optional, Tuple

from bsuite_b200 import registry


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
  """(first global lane, lane count) of `rank`; the remainder goes to the lowest ranks."""
  if not 0 <= rank < world:
    raise ValueError(f'rank {rank} outside world of {world}')
  base, extra = divmod(int(global_batch), int(world))
  count = base + (1 if rank < extra else 0)
  first = rank * base + min(rank, extra)
  return first, count


def load_sharded(bsuite_id: str, global_batch: int, rank: Optional[int] = None, world: Optional[int] = None,
                 device='cuda', seed: int = 0, **engine_kwargs):
  """This rank's shard of a `global_batch`-lane environment (same trajectories as an unsharded one)."""
  import torch.distributed as dist
  if rank is None or world is None:
    if dist.is_available() and dist.is_initialized():
      rank, world = dist.get_rank(), dist.get_world_size()
    else:
      rank, world = 0, 1
  first, count = shard_range(global_batch, rank, world)
  if count == 0:
    raise ValueError(f'rank {rank} would own no lanes (global_batch {global_batch} < world {world})')
  return registry.load_from_id(bsuite_id, batch=count, device=device, seed=seed, lane_offset=first, **engine_kwargs)


def gather_episode_returns(env, group=None) -> Dict[str, Any]:
  """One all-gather of the per-rank reduction of the Logging columns (utils/wrappers.py:113-125).

  Returns tensors of shape [world]: per-rank sums of `steps`, `episode`, `total_return` and the lane count, from
  which the global mean return per episode (the quantity bsuite's analysis consumes) follows.
  """
  import torch
  import torch.distributed as dist
  sums = env.episode_stat_sums()            # device-side reduction: (steps, episode, total_return, len, return)
  block = torch.cat([sums[:3], torch.tensor([float(env.batch)], dtype=torch.float64, device=sums.device)])
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    world = dist.get_world_size(group)
    gathered = torch.empty(world * block.numel(), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(gathered, block, group=group)
    gathered = gathered.view(world, block.numel())
  else:
    gathered = block.view(1, -1)
  return dict(steps=gathered[:, 0], episode=gathered[:, 1], total_return=gathered[:, 2], lanes=gathered[:, 3])


def gather_lane_tensor(tensor, group=None):
  """All-gather of a per-lane tensor [B_rank, ...] into [sum B_rank, ...] (equal shards only)."""
  import torch
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return tensor
  world = dist.get_world_size(group)
  out = torch.empty((world * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
  dist.all_gather_into_tensor(out, tensor.contiguous(), group=group)
  return out
