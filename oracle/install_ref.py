"""Installs the UNMODIFIED reference into git-ignored `oracle/_ref/` so the CPU arm can time its own step() loop.

    python oracle/install_ref.py [--force]

The reference is a pure-Python package; its build writes into the source tree, and /root/reference is read-only,
so a copy under /tmp is handed to pip (offline, no index, no dependency resolution: the four third-party imports
that are absent from this image -- dm_env, immutabledict, termcolor, skimage -- are the stand-ins under
oracle/shims/, none of which contains arithmetic of the step path).  Nothing of the reference enters git history:
`oracle/_ref/` is listed in .gitignore (not in .gpurunignore, so it travels to the GPU box like the built .so).
TEST / MEASUREMENT INFRASTRUCTURE ONLY.
"""

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = '/root/reference'
TARGET = os.path.join(HERE, '_ref')


def installed() -> bool:
  return os.path.isfile(os.path.join(TARGET, 'bsuite', 'environments', 'deep_sea.py'))


def install(force: bool = False) -> str:
  """Returns 'present', 'installed' or 'unavailable: <why>'."""
  if installed() and not force:
    return 'present'
  if not os.path.isdir(os.path.join(REFERENCE, 'bsuite')):
    return f'unavailable: {REFERENCE} does not exist on this machine'
  work = tempfile.mkdtemp(prefix='bsuite_ref_src_')
  try:
    src = os.path.join(work, 'reference')
    shutil.copytree(REFERENCE, src)
    if os.path.isdir(TARGET):
      shutil.rmtree(TARGET)
    cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--quiet',
           '--find-links', '/opt/wheelhouse', '--target', TARGET, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0 or not installed():
      return 'unavailable: pip install failed: ' + (proc.stderr or proc.stdout)[-300:].replace('\n', ' ')
    return 'installed'
  finally:
    shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
  print(install(force='--force' in sys.argv))
