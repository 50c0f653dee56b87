#!/usr/bin/env python
"""GPU-side timeline of the two-phase host step (BSB_HOST_TIMING=1): where one decision's 50-odd microseconds go.

    BSB_HOST_TIMING=1 python tools/e2e_timeline.py [bsuite_id] [batch]

Per step, from %globaltimer stamps the kernel's signaller leaves in the pinned mailbox and from the host's clock:
  gap      previous kernel's last block exit -> this kernel past its dependency wait   (GPU idle between steps)
  phase1   kernel start -> every block has finished its transitions
  fence    phase 1 complete -> scalars fenced to the host (PCIe drain + system fence)
  period   kernel start -> next kernel start
  host     completion word seen -> the next launch call has returned                   (host turnaround)
"""
import ctypes
import os
import sys
import time

os.environ['BSB_HOST_TIMING'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bsuite_b200

BSUITE_ID = sys.argv[1] if len(sys.argv) > 1 else 'deep_sea/11'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
env = bsuite_b200.load_from_id(BSUITE_ID, batch=B, device='cuda', seed=0, track_episodes=True)
ring = [env.make_buffers() for _ in range(4)]
pin = torch.randint(0, env.num_actions, (64, B), dtype=torch.int32).pin_memory()
rows = [pin[i] for i in range(64)]
host = env.make_host_buffers()
lib, handle = env._lib, env._handle.ptr      # pylint: disable=protected-access
stamps = (ctypes.c_uint64 * 8)()
N = 400
rec = np.zeros((N, 4), dtype=np.int64)
host_t = np.zeros((N, 2), dtype=np.float64)
for i in range(30):
  env.step_host(rows[i % 64], host, out=ring[i % 4])
t_prev_done = time.perf_counter()
for i in range(N):
  t0 = time.perf_counter()
  env.step_host(rows[i % 64], host, out=ring[i % 4])
  t1 = time.perf_counter()
  lib.bsb_host_timing(handle, stamps)
  rec[i] = [stamps[0], stamps[1], stamps[2], stamps[3]]
  host_t[i] = [t0, t1]
env.host_flush()
torch.cuda.synchronize()
start, phase1, fenced, prev_exit = rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3]
us = lambda ns: float(np.median(ns)) / 1e3
print(f'{BSUITE_ID} B={B}: medians over {N - 1} steps (us)')
print(f'  period  (start -> next start)            {us(np.diff(start)):7.1f}')
print(f'  gap     (previous exit -> start)         {us((start - prev_exit)[1:]):7.1f}')
print(f'  phase1  (start -> transitions done)      {us(phase1 - start):7.1f}')
print(f'  fence   (transitions done -> fenced)     {us(fenced - phase1):7.1f}')
print(f'  kernel  (start -> last exit, next stamp) {us(prev_exit[1:] - start[:-1]):7.1f}')
print(f'  host: call duration {np.median(host_t[:, 1] - host_t[:, 0]) * 1e6:7.1f}   between calls {np.median(host_t[1:, 0] - host_t[:-1, 1]) * 1e6:7.1f}   call period {np.median(np.diff(host_t[:, 0])) * 1e6:7.1f}')
