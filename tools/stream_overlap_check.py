#!/usr/bin/env python
"""Do fused rollouts of different environments overlap when enqueued on different CUDA streams?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bsuite_b200
from bsuite_b200 import datasets
os.environ[datasets.ENV_VAR] = datasets.write_synthetic_mnist('/tmp/bsb_overlap_mnist', 4096, 16, 0)
ids = ['mnist/0', 'umbrella_distract/22', 'deep_sea/20', 'memory_size/16', 'cartpole/0', 'catch/0']
envs = [bsuite_b200.load_from_id(i, batch=4096, device='cuda', seed=0) for i in ids]
T = 64
bufs = [e.make_buffers(T, with_actions=True) for e in envs]
streams = [torch.cuda.Stream() for _ in envs]

def serial():
  for e, b in zip(envs, bufs):
    e.rollout(T, out=b)

def parallel():
  cur = torch.cuda.current_stream()
  for e, b, s in zip(envs, bufs, streams):
    s.wait_stream(cur)
    with torch.cuda.stream(s):
      e.rollout(T, out=b)
  for s in streams:
    cur.wait_stream(s)

def timed(fn, n=20):
  for _ in range(3): fn()
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): fn()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / n * 1e3

for e, b, i in zip(envs, bufs, ids):
  print(f'{i:24s} alone {timed(lambda: e.rollout(T, out=b)):7.3f} ms per {T}-step rollout')
print(f'all six, one stream      {timed(serial):7.3f} ms')
print(f'all six, six streams     {timed(parallel):7.3f} ms')
