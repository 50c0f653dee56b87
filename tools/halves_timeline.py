#!/usr/bin/env python
"""Where a full step of rollouts.HostHalves goes (BSB_HOST_TIMING=1): host time in submit / collect, the two
half-kernels' %globaltimer intervals and how much of the wall clock the GPU is busy.

    python tools/halves_timeline.py [bsuite_id] [batch] [raw|api] [fence|nofence] [parts]
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
from bsuite_b200 import _lib, rollouts

BSUITE_ID = sys.argv[1] if len(sys.argv) > 1 else 'deep_sea/11'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
RAW = (sys.argv[3] if len(sys.argv) > 3 else 'api') == 'raw'
FENCE = (sys.argv[4] if len(sys.argv) > 4 else 'fence') == 'fence'
PARTS = int(sys.argv[5]) if len(sys.argv) > 5 else 2
halves = rollouts.HostParts(BSUITE_ID, B, device='cuda', seed=0, track_episodes=True, parts=PARTS)
bounds = [0]
for size in halves.sizes:
  bounds.append(bounds[-1] + size)
pin = torch.randint(0, halves.envs[0].num_actions, (64, B), dtype=torch.int32)
rows = [[r for r in pin[:, bounds[h]:bounds[h + 1]].contiguous().pin_memory()] for h in range(PARTS)]
halves.reset()
torch.cuda.synchronize()
lib = halves.envs[0]._lib                                       # pylint: disable=protected-access
handles = [e._handle.ptr for e in halves.envs]                   # pylint: disable=protected-access
houts = [h.as_outputs() for h in halves.host]
dev_obs = [o.observation.data_ptr() for o in halves.out]
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flags = _lib.HOST_NO_WAIT | (_lib.HOST_FENCE_CALLER if FENCE else 0)
stamps = (ctypes.c_uint64 * 8)()
N = 400
rec = np.zeros((N, PARTS, 4), dtype=np.int64)
host_t = np.zeros((N, PARTS, 4), dtype=np.float64)      # collect start, collect end, submit start, submit end


def submit(h, t):
  if RAW:
    rc = lib.bsb_step_host(handles[h], rows[h][t % 64].data_ptr(), ctypes.byref(houts[h]), dev_obs[h], stream, flags)
    assert rc == 0
  else:
    halves.submit(h, rows[h][t % 64])


def collect(h):
  if RAW:
    assert lib.bsb_host_wait(handles[h]) == 0
  else:
    halves.collect(h)


for h in range(PARTS):
  submit(h, 0)
for t in range(1, 30):
  for h in range(PARTS):
    collect(h); submit(h, t)
for t in range(N):
  for h in range(PARTS):
    c0 = time.perf_counter(); collect(h); c1 = time.perf_counter()
    lib.bsb_host_timing(handles[h], stamps)
    rec[t, h] = [stamps[0], stamps[1], stamps[2], stamps[3]]
    s0 = time.perf_counter(); submit(h, t); s1 = time.perf_counter()
    host_t[t, h] = [c0, c1, s0, s1]
for h in range(PARTS):
  collect(h)
torch.cuda.synchronize()
us = lambda x: float(np.median(x)) / 1e3
start, phase1, fenced, prev_exit = rec[..., 0], rec[..., 1], rec[..., 2], rec[..., 3]
print(f'{BSUITE_ID} B={B} {PARTS} parts ({"raw ctypes" if RAW else "HostHalves API"}, {"fence" if FENCE else "no fence"}): medians over {N} steps (us)')
print(f'  full-step period (half 0 start -> next start)   {us(np.diff(start[:, 0])):7.1f}')
for h in range(PARTS):
  print(f'  half {h}: phase1 {us(phase1[:, h] - start[:, h]):6.1f}  fence {us(fenced[:, h] - phase1[:, h]):6.1f}  '
        f'kernel (start -> last exit) {us(prev_exit[1:, h] - start[:-1, h]):6.1f}  '
        f'own gap (exit -> next start) {us(start[1:, h] - prev_exit[1:, h]):6.1f}')
# GPU busy: union of the half-kernel intervals [start(t), exit(t)] (exit(t) is stamped into step t+1)
iv = sorted([(int(start[t, h]), int(prev_exit[t + 1, h])) for t in range(N - 1) for h in range(PARTS)])
busy, cur_s, cur_e = 0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
  if s > cur_e:
    busy += cur_e - cur_s; cur_s, cur_e = s, e
  else:
    cur_e = max(cur_e, e)
busy += cur_e - cur_s
span = iv[-1][1] - iv[0][0]
print(f'  GPU busy {busy / span:6.3f} of the span; idle per full step {(span - busy) / (N - 1) / 1e3:6.1f}; '
      f'start(part 1) - exit(part 0) {us(start[:-1, 1] - prev_exit[1:, 0]):6.1f}; start(part 0, t+1) - exit(last part, t) {us(start[1:-1, 0] - prev_exit[2:, PARTS - 1]):6.1f}')
print(f'  host per half-step: collect {np.median(host_t[..., 1] - host_t[..., 0]) * 1e6:6.1f}  submit {np.median(host_t[..., 3] - host_t[..., 2]) * 1e6:6.1f}; '
      f'collects that returned within 1 us: {np.mean((host_t[..., 1] - host_t[..., 0]) < 1.5e-6):5.2f}; '
      f'host period per full step {np.median(np.diff(host_t[:, 0, 0])) * 1e6:6.1f}')
halves.close()
