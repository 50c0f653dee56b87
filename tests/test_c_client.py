"""The C ABI from plain C: compiles tests/c_client/deep_sea_client.c against include/bsuite_b200.h (strict C99,
warnings are errors), runs BASELINE config #1 on the explicit host path and compares the program's output with the
known answers recorded from the unmodified reference (deep_sea/0, 91 episodes)."""

import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from bsuite_b200 import _lib
from tests import conftest as cf

CLIENT = os.path.join(cf.ROOT, 'tests', 'c_client', 'deep_sea_client.c')


@pytest.mark.skipif(shutil.which('gcc') is None, reason='needs gcc')
def test_c99_client_reproduces_config_1(tmp_path):
  _lib.load()
  lib_dir = os.path.dirname(_lib.LIB_PATH)
  binary = str(tmp_path / 'deep_sea_client')
  subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Wextra', '-Werror', '-O1',
                  '-I', os.path.join(cf.ROOT, 'include'), CLIENT, '-o', binary,
                  '-L', lib_dir, '-lbsuite_b200', '-Wl,-rpath,' + lib_dir], check=True, capture_output=True, text=True)
  size, count = 10, 1000
  # the SAME numpy calls the reference makes: deep_sea.py:80-81 (mapping_seed = 42 in every sweep entry)
  mapping = np.random.RandomState(42).binomial(1, 0.5, [size, size]).astype(np.uint8)
  actions = np.random.RandomState(0).randint(2, size=count).astype(np.int32)
  mapping.tofile(tmp_path / 'mapping.u8')
  actions.tofile(tmp_path / 'actions.i32')
  proc = subprocess.run([binary, str(size), str(tmp_path / 'mapping.u8'), str(tmp_path / 'actions.i32'), str(count)],
                        capture_output=True, text=True)
  assert proc.returncode == 0, proc.stderr
  got = {}
  for line in proc.stdout.splitlines():
    fields = line.split()
    if fields[0] == 'info':
      fields = fields[1:]
    got[fields[0]] = float(fields[1])
  row = next(r for r in json.load(open(os.path.join(cf.GOLDEN_DIR, 'known_answers.json'))) if r['label'] == 'deep_sea/0')
  assert got['abi'] == _lib.ABI_VERSION
  assert got['num_last'] == row['num_last'] == 91
  assert got['num_first'] == 91                       # reset() + an auto-reset after each LAST but the final one
  assert got['reward_sum'] == pytest.approx(row['reward_sum'], abs=1e-12)
  assert {k: got[k] for k in row['info']} == row['info']
  assert got['hot_cells'] == count - 91                # one hot cell per observation, none on LAST (deep_sea.py:137-139)
  assert got['steps_done'] == count + 1
  assert got['null_handle_status'] == 1                # BSB_INVALID_ARGUMENT
