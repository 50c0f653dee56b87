"""INTEGRATION.md's reference-side binding, executed: `examples/reference_binding_stub.py` (plain ctypes, nothing
from the bsuite_b200 package) reproduces the unmodified reference's deep_sea known answers -- digest, #LAST,
bsuite_info -- and plugs into the reference's own registry when the reference is importable."""

import hashlib
import importlib.util
import json
import os
import struct

import numpy as np
import pytest

from oracle import reference_runner as rr
from tests import conftest as cf


def _stub():
  path = os.path.join(cf.ROOT, 'examples', 'reference_binding_stub.py')
  spec = importlib.util.spec_from_file_location('reference_binding_stub', path)
  module = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(module)
  return module


def _digest(rows):
  h = hashlib.sha256()
  for ts in rows:
    h.update(struct.pack('<i', int(ts.step_type)))
    h.update(struct.pack('<d', float('nan') if ts.reward is None else float(ts.reward)))
    h.update(struct.pack('<d', float('nan') if ts.discount is None else float(ts.discount)))
    h.update(np.ascontiguousarray(ts.observation, dtype=np.float32).tobytes())
  return h.hexdigest()[:16]


def _rows():
  answers = json.load(open(os.path.join(cf.GOLDEN_DIR, 'known_answers.json')))
  return [r for r in answers if r['label'] in ('deep_sea/0', 'deep_sea/11')]


@pytest.mark.parametrize('row', _rows(), ids=lambda r: r['label'])
def test_stub_reproduces_reference_known_answers(row):
  size = {'deep_sea/0': 10, 'deep_sea/11': 32}[row['label']]
  env = _stub().DeepSeaB200(size=size, mapping_seed=42)        # deep_sea/sweep.py:20
  actions = np.random.RandomState(0).randint(2, size=1000)
  rows = [env.reset()] + [env.step(int(a)) for a in actions]
  assert rows[0].first() and rows[0].reward is None and rows[0].discount is None
  assert sum(ts.last() for ts in rows) == row['num_last']
  assert _digest(rows) == row['digest']
  assert {k: float(v) for k, v in env.bsuite_info().items()} == row['info']
  env.close()


@pytest.mark.skipif(not rr.reference_available(), reason='the reference tree is not present on this machine')
def test_stub_registers_in_the_reference():
  """bsuite.load_from_id('deep_sea/0') through the reference's OWN registry and sweep tables, with the stub class
  swapped in for the numpy implementation: same trace as the reference's class."""
  rr.import_reference()
  import bsuite  # pylint: disable=import-outside-toplevel
  from bsuite import bsuite as bsuite_registry  # pylint: disable=import-outside-toplevel
  stub = _stub()
  original = bsuite_registry.EXPERIMENT_NAME_TO_ENVIRONMENT['deep_sea']
  actions = np.random.RandomState(3).randint(2, size=300)
  want_env = bsuite.load_from_id('deep_sea/2')
  want = [want_env.reset()] + [want_env.step(int(a)) for a in actions]
  try:
    bsuite_registry.EXPERIMENT_NAME_TO_ENVIRONMENT['deep_sea'] = (
        lambda **kwargs: stub.DeepSeaB200(**kwargs))
    env = bsuite.load_from_id('deep_sea/2')
    assert isinstance(env, stub.DeepSeaB200)
    got = [env.reset()] + [env.step(int(a)) for a in actions]
  finally:
    bsuite_registry.EXPERIMENT_NAME_TO_ENVIRONMENT['deep_sea'] = original
  assert _digest(got) == _digest(want)
