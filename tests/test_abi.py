"""The C ABI: every symbol include/bsuite_b200.h declares is exported and callable; errors are loud."""

import ctypes
import os
import re

import numpy as np
import pytest

from bsuite_b200 import _lib
from tests import conftest as cf

HEADER = os.path.join(cf.ROOT, 'include', 'bsuite_b200.h')


def _declared_functions():
  text = open(HEADER).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(bsb_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
  declared = _declared_functions()
  assert declared, 'no functions parsed from the header'
  assert sorted(_lib.EXPORTS) == declared


def test_library_exports_every_declared_symbol():
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in _declared_functions():
    assert hasattr(lib, name), f'{name} is declared in include/bsuite_b200.h but not exported'
  assert _lib.load().bsb_abi_version() == _lib.ABI_VERSION
  header_version = int(re.search(r'#define BSB_ABI_VERSION (\d+)', open(HEADER).read()).group(1))
  assert header_version == _lib.ABI_VERSION


def test_library_has_no_torch_dependency():
  """The boundary is plain C: the .so must not link against torch / python."""
  import subprocess
  out = subprocess.run(['ldd', _lib.LIB_PATH], capture_output=True, text=True).stdout
  assert 'torch' not in out and 'python' not in out, out


def test_sm100a_code_is_embedded():
  import shutil
  import subprocess
  cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
  if not os.path.exists(cuobjdump):
    pytest.skip('cuobjdump not available')
  out = subprocess.run([cuobjdump, '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
  assert 'sm_100a' in out, out


def test_invalid_arguments_return_status_and_message():
  lib = _lib.load()
  cfg = _lib.Config()
  cfg.family = 99
  handle = ctypes.c_void_p()
  status = lib.bsb_create(ctypes.byref(cfg), 4, _lib.DEVICE_HOST, 0, 0, ctypes.byref(handle))
  assert status == 1 and b'unknown family' in lib.bsb_last_error()
  cfg.family = _lib.DEEP_SEA
  cfg.size = 10                      # mapping table missing
  status = lib.bsb_create(ctypes.byref(cfg), 4, _lib.DEVICE_HOST, 0, 0, ctypes.byref(handle))
  assert status == 1 and b'mapping' in lib.bsb_last_error()
  cfg.family = _lib.MEMORY_CHAIN
  cfg.memory_length, cfg.num_bits = 3, 65
  status = lib.bsb_create(ctypes.byref(cfg), 4, _lib.DEVICE_HOST, 0, 0, ctypes.byref(handle))
  assert status == 2 and b'num_bits' in lib.bsb_last_error()
  cfg.num_bits = 3
  status = lib.bsb_create(ctypes.byref(cfg), 0, _lib.DEVICE_HOST, 0, 0, ctypes.byref(handle))
  assert status == 1 and b'batch' in lib.bsb_last_error()
  with pytest.raises(_lib.EngineError):
    _lib.check(status)


def test_cuda_request_without_device_fails_loudly():
  """No implicit CPU fallback: asking for a CUDA device on a machine without one raises."""
  import torch
  import bsuite_b200
  if torch.cuda.is_available():
    pytest.skip('a CUDA device is present')
  with pytest.raises(RuntimeError, match='no implicit CPU fallback'):
    bsuite_b200.load_from_id('catch/0', batch=8)           # default device is cuda
  with pytest.raises(RuntimeError, match='no implicit CPU fallback'):
    bsuite_b200.load_from_id('catch/0')                    # B = 1 adapter too
  lib = _lib.load()
  cfg = _lib.Config()
  cfg.family, cfg.rows, cfg.columns = _lib.CATCH, 10, 5
  handle = ctypes.c_void_p()
  status = lib.bsb_create(ctypes.byref(cfg), 8, 0, 0, 0, ctypes.byref(handle))   # device ordinal 0
  assert status == 3 and b'no CUDA device' in lib.bsb_last_error()


def test_raw_abi_round_trip_without_torch():
  """Drives the host path with nothing but ctypes + numpy, as a foreign-language binding would."""
  lib = _lib.load()
  cfg = _lib.Config()
  cfg.family, cfg.rows, cfg.columns, cfg.reward_scale = _lib.CATCH, 10, 5, 1.0
  handle = ctypes.c_void_p()
  _lib.check(lib.bsb_create(ctypes.byref(cfg), 3, _lib.DEVICE_HOST, 5, 100, ctypes.byref(handle)))
  numel = ctypes.c_int64()
  _lib.check(lib.bsb_obs_numel(handle, ctypes.byref(numel)))
  rows, cols, n_act = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
  _lib.check(lib.bsb_obs_shape(handle, ctypes.byref(rows), ctypes.byref(cols)))
  _lib.check(lib.bsb_num_actions(handle, ctypes.byref(n_act)))
  assert (numel.value, rows.value, cols.value, n_act.value) == (50, 10, 5, 3)
  obs = np.zeros((3, 50), np.float32)
  reward = np.zeros(3, np.float32)
  step_type = np.full(3, -1, np.int32)
  out = _lib.Outputs()
  out.observation, out.reward, out.step_type = obs.ctypes.data, reward.ctypes.data, step_type.ctypes.data
  actions = np.array([0, 1, 2], np.int32)
  for t in range(10):
    _lib.check(lib.bsb_step(handle, ctypes.c_void_p(actions.ctypes.data), ctypes.byref(out), None))
    assert list(step_type) == [0 if t == 0 else (2 if t == 9 else 1)] * 3
    assert np.all(obs.sum(axis=1) >= 1.0) and np.all(obs.sum(axis=1) <= 2.0)
  steps = ctypes.c_int64()
  _lib.check(lib.bsb_steps_done(handle, ctypes.byref(steps)))
  assert steps.value == 10
  count = ctypes.c_int32()
  _lib.check(lib.bsb_info_count(handle, ctypes.byref(count)))
  assert count.value == 1 and lib.bsb_info_name(handle, 0) == b'total_regret'
  regret = np.zeros(3, np.float64)
  _lib.check(lib.bsb_read_info(handle, 0, ctypes.c_void_p(regret.ctypes.data), None))
  assert set(regret.tolist()) <= {0.0, 2.0}
  assert lib.bsb_read_info(handle, 5, ctypes.c_void_p(regret.ctypes.data), None) == 1
  _lib.check(lib.bsb_destroy(handle))


def test_abi_fuzzer_finds_no_crash():
  """A short run of tools/fuzz_abi.py (hostile configurations and argument misuse on the host path); the long run
  under ASan/UBSan is tools/host_sanitize.sh."""
  import subprocess
  import sys
  proc = subprocess.run([sys.executable, os.path.join(cf.ROOT, 'tools', 'fuzz_abi.py'), '300', '7'],
                        capture_output=True, text=True, timeout=300)
  assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
  assert 'no crash' in proc.stdout


def test_communicator_entry_points_validate_their_arguments():
  """bsb_comm_* (NCCL resolved at run time): the id can be drawn without a GPU; creation needs a CUDA device."""
  import ctypes
  lib = _lib.load()
  buf = (ctypes.c_uint8 * _lib.COMM_ID_BYTES)()
  status = lib.bsb_comm_unique_id(buf)
  if status != 0:                         # no NCCL library on this machine: reported, not crashed
    assert b'NCCL' in lib.bsb_last_error()
    return
  assert any(bytes(buf))
  handle = ctypes.c_void_p()
  assert lib.bsb_comm_create(buf, 3, 2, 0, ctypes.byref(handle)) != 0 and not handle.value      # rank >= world
  assert lib.bsb_comm_create(None, 0, 1, 0, ctypes.byref(handle)) != 0
  assert lib.bsb_comm_destroy(None) == 0 and lib.bsb_comm_wait(None, None) != 0
