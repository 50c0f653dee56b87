#!/usr/bin/env python
"""Which members bound the fused lock-step of BASELINE config #5?  Times 64-step fused rollouts (one graph per
subset) of the whole 23-experiment batch, of each experiment alone and of a few subsets, at `lanes` lanes per id.

    python tools/sweep_breakdown.py [lanes]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bsuite_b200 import datasets, suite

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = 64
os.environ[datasets.ENV_VAR] = datasets.write_synthetic_mnist('/tmp/bsb_sweep_breakdown_mnist', 4096, 16, 0)


def timed(ids):
  batch = suite.SweepBatch(ids, lanes=lanes, device='cuda', seed=0)
  graphed = batch.capture(T)
  for _ in range(3):
    graphed.replay()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    graphed.replay()
  e1.record()
  torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 1e3 / (10 * T)
  del graphed
  batch.close()
  return us


ids = suite.one_per_experiment()
print(f'{lanes} lanes per id, {T}-step fused rollouts replayed from one graph: microseconds per lock-step')
print(f'  all {len(ids)} ids together            {timed(ids):7.2f}')
alone = {i: timed([i]) for i in ids}
for i, us in sorted(alone.items(), key=lambda kv: -kv[1]):
  print(f'  {i:28s} alone  {us:7.2f}')
print(f'  sum of the ids alone             {sum(alone.values()):7.2f}     slowest alone {max(alone.values()):7.2f}')
groups = {'mnist*': [i for i in ids if i.startswith('mnist')], 'cartpole* + mountain_car*': [i for i in ids if i.startswith(('cartpole', 'mountain'))],
          'everything but mnist*': [i for i in ids if not i.startswith('mnist')]}
for name, members in groups.items():
  print(f'  {name:34s} {timed(members):7.2f}')
