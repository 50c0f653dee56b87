#!/usr/bin/env python
"""Where the time of one host-driven step goes (deep_sea N=32, B=65536, one GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bsuite_b200

env = bsuite_b200.load_from_id('deep_sea/11', batch=65536, device='cuda', seed=0, track_episodes=True)
ring = [env.make_buffers() for _ in range(4)]
dev_acts = torch.randint(0, 2, (64, 65536), device='cuda', dtype=torch.int32)
rows = [dev_acts[i] for i in range(64)]
pin = torch.randint(0, 2, (64, 65536), dtype=torch.int32).pin_memory()
prow = [pin[i] for i in range(64)]
host = env.make_host_buffers()

def timed(fn, n=300):
  for i in range(20): fn(i)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for i in range(n): fn(i)
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / n * 1e6

def dev_async(i): env.step(rows[i % 64], out=ring[i % 4])
def dev_sync(i): env.step(rows[i % 64], out=ring[i % 4]); torch.cuda.synchronize()
def host_zc(i): env.step_host(prow[i % 64], host, out=ring[i % 4])
print(f'device actions, launches queued (kernel rate)      {timed(dev_async):7.1f} us/step')
print(f'device actions, synchronise after every step       {timed(dev_sync):7.1f} us/step')
print(f'pinned host actions + host scalars (zero-copy)     {timed(host_zc):7.1f} us/step')
os.environ['BSB_ZERO_COPY'] = '0'
env2 = bsuite_b200.load_from_id('deep_sea/11', batch=65536, device='cuda', seed=0, track_episodes=True)
def host_staged(i): env2.step_host(prow[i % 64], host, out=ring[i % 4])
print(f'pinned host actions + host scalars (staged copies) {timed(host_staged):7.1f} us/step')
