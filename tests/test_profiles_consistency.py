"""The bench lines committed under profiles/ are internally consistent: every derived figure follows from the
measured ones on the same line (the recomputation a reader of the evidence would do)."""

import glob
import json
import os

import pytest

from tests import conftest as cf

LINES = sorted(glob.glob(os.path.join(cf.ROOT, 'profiles', 'r02*bench*n1*.json')) +
               glob.glob(os.path.join(cf.ROOT, 'profiles', 'r02_scale8_n[1248].json')) +
               glob.glob(os.path.join(cf.ROOT, 'profiles', 'r02_last_ab_*.json')))


def _engine_line(path):
  for raw in open(path).read().strip().splitlines():
    if raw.startswith('{'):
      line = json.loads(raw)
      if line.get('impl', 'b200') != 'reference' and 'roofline' in line:
        return line
  return None


@pytest.mark.parametrize('path', LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_line_is_consistent(path):
  line = _engine_line(path)
  if line is None:
    pytest.skip('no engine line in this file')
  roof = line['roofline']
  lanes = 65536 * line['n_gpus']
  # value = lanes x steps / median window; ms_per_step = that window / steps
  assert line['value'] == pytest.approx(lanes / (line['ms_per_step'] * 1e-3), rel=1e-6)
  if 'windows_ms' in line:
    windows = sorted(line['windows_ms'])
    assert windows[len(windows) // 2] / line['steps'] == pytest.approx(line['ms_per_step'], rel=1e-6)
  # roofline: achieved = algorithmic bytes per launch / launch time; frac = achieved / peak; the window can only be slower
  assert roof['algorithmic_bytes_per_launch'] == 4120 * 65536
  assert roof['achieved'] == pytest.approx(roof['algorithmic_bytes_per_launch'] / (roof['launch_us'] * 1e-6) / 1e9, rel=1e-6)
  assert roof['frac'] == pytest.approx(roof['achieved'] / roof['peak'], rel=1e-9) and roof['frac'] <= 1.02
  assert roof['frac_from_ms_per_step'] <= roof['frac'] + 1e-9
  assert roof['launch_us'] * 1e-3 <= line['ms_per_step'] * (1 + 1e-6)
  if roof.get('traffic') is not None:      # DRAM traffic never exceeds the algorithmic bytes by more than a few per cent
    assert roof['traffic'] <= 1.05 * roof['algorithmic_bytes_per_launch']
  # e2e: the strict host loop cannot beat the device-resident one; the headline is the fastest mode reported
  e2e = line['e2e']
  assert 0 < e2e['value'] < line['value']
  assert e2e['h2d_bytes_per_step'] == 4 * 65536 and e2e['d2h_bytes_per_step'] == 12 * 65536
  modes = [e2e.get('one_batch_value')] + list((e2e.get('parts_values') or {}).values()) + [e2e.get('two_halves_value')]
  modes = [m for m in modes if m]
  if modes:
    assert e2e['value'] == pytest.approx(max(modes), rel=1e-9)
  clocks = line.get('clocks')
  if clocks:
    assert not set(clocks['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    assert clocks['sm_mhz'] >= 0.9 * clocks['sm_max_mhz']
