#!/usr/bin/env python
"""BASELINE config #5: the 23-experiment heterogeneous batch, `lanes` lanes per experiment, sharded over the ranks.

    python tools/bench_sweep.py [--lanes 4096] [--steps 64] [--iters 20]
    torchrun --nproc-per-node 8 tools/bench_sweep.py --gpus 8

Every rank owns lanes/W lanes of EVERY experiment; one iteration = one fused `steps`-step rollout of all 23
environments (each on its own CUDA stream) + one NCCL all-gather of the per-id return statistics.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bsuite_b200 import datasets  # noqa: E402
from bsuite_b200 import suite  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--lanes', type=int, default=4096)
  ap.add_argument('--steps', type=int, default=64)
  ap.add_argument('--iters', type=int, default=20)
  args = ap.parse_args()
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  local_rank = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)
  mnist_dir = f'/tmp/bsb_sweep_mnist_{rank}'
  datasets.write_synthetic_mnist(mnist_dir, 4096, 16, 0)
  os.environ[datasets.ENV_VAR] = mnist_dir
  batch = suite.SweepBatch(lanes=args.lanes, device=device, seed=0, rank=rank, world=world)
  for _ in range(3):
    batch.rollout(args.steps)
    batch.gather_returns()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.iters):
    batch.rollout(args.steps)
    gathered = batch.gather_returns()
  e1.record()
  torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  seconds = float(ms[0]) * 1e-3
  lane_steps = len(batch.bsuite_ids) * args.lanes * args.steps * args.iters
  if rank == 0:
    totals = gathered.sum(dim=0).cpu()
    print(json.dumps({
        'workload': f'{len(batch.bsuite_ids)} experiments x {args.lanes} lanes, {args.steps}-step fused rollouts, '
                    f'sharded over {world} GPU(s)',
        'n_gpus': world, 'env_steps_per_s': lane_steps / seconds, 'us_per_lockstep': seconds / (args.steps * args.iters) * 1e6,
        'algorithmic_gbs': world * batch.bytes_per_step() * args.steps * args.iters / seconds / 1e9,
        'episodes_total': float(totals[:, 1].sum()), 'steps_total': float(totals[:, 2].sum()),
        'per_id_mean_return_per_episode': {k: float(totals[i, 0] / max(float(totals[i, 1]), 1.0)) for i, k in enumerate(batch.bsuite_ids)},
    }))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
