#!/usr/bin/env python
"""Multi-rank checks of the one collective of the path (run under torchrun, one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

  1. lanes sharded over the ranks reproduce the unsharded trajectories (rank 0 steps the whole batch as well);
  2. the asynchronous log point (`distributed.LogPoint`: reduction in stream order, NCCL all-gather on a side
     stream) equals the synchronous gather, with steps queued behind it;
  3. the C ABI's own communicator (`bsb_comm_*` / `bsb_log_point`, NCCL loaded by the library) equals both.
Prints one JSON line on rank 0; exits non-zero on any mismatch.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bsuite_b200  # noqa: E402
from bsuite_b200 import distributed as bd  # noqa: E402


def main():
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  dist.init_process_group('nccl', device_id=device)
  global_batch, T = 8192, 40
  ids = ['catch/0', 'deep_sea/0', 'cartpole/0']
  envs = [bd.load_sharded(i, global_batch, rank=rank, world=world, device=device, seed=4, track_episodes=True) for i in ids]
  async_lp = bd.LogPoint(envs, slots=2)
  native_lp = bd.NativeLogPoint(envs)
  ok = True
  for round_ in range(4):
    for env in envs:
      env.rollout(T, action_seed=round_)
    sync = torch.stack([torch.stack([bd.gather_episode_returns(env)[k] for k in ('steps', 'episode', 'total_return')], dim=-1)
                        for env in envs], dim=1)                    # [world, n_envs, 3]
    ticket = async_lp.issue()
    native_lp.issue()
    for env in envs:                                                # work queued behind both log points
      env.rollout(3, action_seed=100 + round_)
    got_async = async_lp.result(ticket)[..., :3]
    got_native = native_lp.result()[..., :3]
    torch.cuda.synchronize()
    a_ok, n_ok = torch.equal(got_async, sync), torch.equal(got_native, sync)
    if not (a_ok and n_ok):
      print(f'[rank {rank}] round {round_}: async==sync {a_ok}, native==sync {n_ok}\n sync {sync.tolist()}\n async {got_async.tolist()}\n native {got_native.tolist()}', flush=True)
    ok = ok and a_ok and n_ok
  # sharding invariance: rank 0 also runs the whole batch and compares its own slice and the gathered totals
  totals = async_lp.result(async_lp.issue(), host_sync=True).sum(dim=0)     # [n_envs, 5]
  if rank == 0:
    for k, bsuite_id in enumerate(ids):
      whole = bsuite_b200.load_from_id(bsuite_id, batch=global_batch, device=device, seed=4, track_episodes=True)
      for round_ in range(4):
        whole.rollout(T, action_seed=round_)
        whole.rollout(3, action_seed=100 + round_)
      want = whole.episode_stat_sums()
      # sums over lanes of float rewards depend on the partition in the last bits; the per-lane values do not
      same = torch.allclose(totals[k], want, rtol=1e-12, atol=1e-6 if bsuite_id.startswith('cartpole') else 1e-9)
      if not same:
        print(f'[rank 0] sharding: {bsuite_id}: sharded totals {totals[k].tolist()} vs whole {want.tolist()}', flush=True)
      ok = ok and bool(same)
      whole.close()
  flag = torch.tensor([1.0 if ok else 0.0], device=device)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    print(json.dumps({'world': world, 'async_equals_sync': bool(flag.item() == 1.0), 'ids': ids, 'global_batch': global_batch}))
  native_lp.close()
  dist.destroy_process_group()
  return 0 if flag.item() == 1.0 else 1


if __name__ == '__main__':
  sys.exit(main())
