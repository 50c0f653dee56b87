#!/usr/bin/env python
"""Where the time of one host-driven step goes (deep_sea N=32, B=65536, one GPU).

    python tools/e2e_breakdown.py [bsuite_id] [batch]

Rows: the kernel rate (launches queued), a device-resident loop that synchronises after every step, then the
host-buffer call `BatchedEnvironment.step_host` (pinned actions in, pinned scalars out) in its variants:
stream synchronise (BSB_HOST_SPIN=0), mailbox completion, two-phase (scalars first), pre-launched doorbell
kernels -- each through the Python face and through bare ctypes calls with prebuilt arguments (what a C caller
of the ABI pays).
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bsuite_b200
from bsuite_b200 import _lib

BSUITE_ID = sys.argv[1] if len(sys.argv) > 1 else 'deep_sea/11'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536


def make(**env_vars):
  for k, v in env_vars.items():
    os.environ[k] = v
  env = bsuite_b200.load_from_id(BSUITE_ID, batch=B, device='cuda', seed=0, track_episodes=True)
  for k in env_vars:
    os.environ.pop(k)
  return env


def timed(fn, n=300, after=None):
  for i in range(20):
    fn(i)
  if after:
    after()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(n):
    fn(i)
  if after:
    after()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / n * 1e6


env = make()
ring = [env.make_buffers() for _ in range(4)]
n_act = env.num_actions
dev_acts = torch.randint(0, n_act, (64, B), device='cuda', dtype=torch.int32)
rows = [dev_acts[i] for i in range(64)]
pin = torch.randint(0, n_act, (64, B), dtype=torch.int32).pin_memory()
prow = [pin[i] for i in range(64)]
host = env.make_host_buffers()

print(f'{BSUITE_ID} B={B}')
print(f'device actions, launches queued (kernel rate)            {timed(lambda i: env.step(rows[i % 64], out=ring[i % 4])):7.1f} us/step')


def dev_sync(i):
  env.step(rows[i % 64], out=ring[i % 4])
  torch.cuda.synchronize()


print(f'device actions, synchronise after every step             {timed(dev_sync):7.1f} us/step')


def variants(e, label):
  lib, handle = e._lib, e._handle.ptr          # pylint: disable=protected-access
  houts = host.as_outputs()
  acts = [ctypes.c_void_p(p.data_ptr()) for p in prow]
  obs = [ctypes.c_void_p(r.observation.data_ptr()) for r in ring]
  ref = ctypes.byref(houts)
  for prelaunch in (False, True):
    flags = _lib.HOST_PRELAUNCH if prelaunch else 0
    py = timed(lambda i: e.step_host(prow[i % 64], host, out=ring[i % 4], prelaunch=prelaunch), after=e.host_flush)
    raw = timed(lambda i: lib.bsb_step_host(handle, acts[i % 64], ref, obs[i % 4], None, flags), after=e.host_flush)
    print(f'step_host {label:34s} prelaunch={int(prelaunch)}  python {py:6.1f}  ctypes {raw:6.1f} us/step')


variants(make(BSB_HOST_SPIN='0'), 'stream synchronise (round 1)')
variants(make(BSB_HOST_EARLY='0'), 'mailbox completion')
variants(make(), 'mailbox + two-phase (default)')
os.environ['BSB_ZERO_COPY'] = '0'
env2 = bsuite_b200.load_from_id(BSUITE_ID, batch=B, device='cuda', seed=0, track_episodes=True)
print(f'step_host staged copies (BSB_ZERO_COPY=0)                {timed(lambda i: env2.step_host(prow[i % 64], host, out=ring[i % 4])):7.1f} us/step')
