#!/usr/bin/env python
"""A few round-robin steps of rollouts.HostParts for an `ncu` launch list of the split host step's two launches.

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:transition_kernel --launch-skip 12 \
        --launch-count 16 --csv --log-file gpurun_out/split_launches.csv python tools/parts_probe.py [parts]

Under ncu the launches are serialised, so the durations are those of each launch ALONE: the transitions + copiers
launch (grid of 128-thread CTAs, no shared memory) and the observation-only launch (persistent 32-thread CTAs).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bsuite_b200 import rollouts

PARTS = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, T = 65536, 8
group = rollouts.HostParts('deep_sea/11', B, device='cuda', seed=0, track_episodes=True, parts=PARTS)
bounds = [0]
for size in group.sizes:
  bounds.append(bounds[-1] + size)
pin = torch.randint(0, 2, (T, B), dtype=torch.int32)
rows = [[r for r in pin[:, bounds[p]:bounds[p + 1]].contiguous().pin_memory()] for p in range(PARTS)]
group.reset()
for p in range(PARTS):
  group.submit(p, rows[p][0])
for t in range(1, T):
  for p in range(PARTS):
    group.collect(p); group.submit(p, rows[p][t])
for p in range(PARTS):
  group.collect(p)
torch.cuda.synchronize()
group.close()
print('parts_probe done', PARTS, group.sizes)
