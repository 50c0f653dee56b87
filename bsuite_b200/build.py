"""Builds libbsuite_b200.so in-tree with nvcc for sm_100a.

    python -m bsuite_b200.build [--force]

The library is the only compiled artefact: CUDA kernels, the explicit host path
and the extern "C" surface of include/bsuite_b200.h.  It links cudart statically
and depends on nothing from torch.
"""

import os
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SOURCES = [os.path.join(CSRC, 'bsb_engine.cu')]
HEADERS = [os.path.join(CSRC, f) for f in ('bsb_rng.cuh', 'bsb_families.cuh', 'bsb_kernels.cuh')] + [
    os.path.join(os.path.dirname(HERE), 'include', 'bsuite_b200.h')]
OUTPUT = os.path.join(HERE, 'libbsuite_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '--fmad=false',                       # CPython/numpy never fuse a*b+c (float-dynamics parity)
    '-Xcompiler', '-fPIC,-ffp-contract=off,-O2',
    '-shared',
]


def find_nvcc() -> str:
  for candidate in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
    if candidate and os.path.exists(candidate):
      return candidate
  raise RuntimeError('nvcc not found: bsuite_b200 needs the CUDA toolkit to build (no CPU-only build exists)')


def is_stale() -> bool:
  if not os.path.exists(OUTPUT):
    return True
  built = os.path.getmtime(OUTPUT)
  return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
  if not force and not is_stale():
    return OUTPUT
  cmd = [find_nvcc()] + NVCC_FLAGS + ['-o', OUTPUT] + SOURCES
  if verbose:
    cmd += ['-Xptxas', '-v']
  start = time.time()
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
  if verbose:
    sys.stderr.write(proc.stderr)
  sys.stderr.write(f'[bsuite_b200.build] built {OUTPUT} in {time.time() - start:.1f}s\n')
  return OUTPUT


if __name__ == '__main__':
  build_library(force='--force' in sys.argv, verbose='-v' in sys.argv)
