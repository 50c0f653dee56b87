"""Parity of the engine against traces recorded from the UNMODIFIED reference.

Every fixture under tests/golden/ (made by oracle/gen_golden.py) holds, for 8
lanes and T calls, the reference's step_type / reward / discount / observation
and final bsuite_info().  The engine is driven through the C ABI with the same
actions; integer/grid families must match BIT-EXACTLY, float-dynamics families
within 1e-6 (north_star).  The `cpu` variant exercises the explicit host path of
the C ABI (BASELINE config #1's "no GPU" plumbing); the `cuda` variant is the
product path and runs on the B200 box (-m gpu).
"""

import numpy as np
import pytest

import bsuite_b200
from tests import conftest as cf


def _run_engine(meta, data, device, fused):
  import torch
  kwargs = dict(meta['kwargs'])
  wrap = {}
  if meta['wrapper'] == 'noise':
    wrap['noise_scale'] = meta['wrapper_arg']
  elif meta['wrapper'] == 'scale':
    wrap['reward_scale'] = meta['wrapper_arg']
  env = bsuite_b200.make(meta['env_class'], batch=len(meta['lanes']), device=device, seed=meta['seed'],
                         rng=meta['rng'], engine_kwargs=dict(reward_dtype='float64'), **wrap, **kwargs)
  actions = torch.as_tensor(data['actions'])
  T = actions.shape[0]
  reset_at = set(meta['reset_at'])
  out = dict(step_type=[], reward=[], discount=[], observation=[])

  def take(ts):
    for k in out:
      out[k].append(getattr(ts, k).cpu().numpy().copy())

  if fused and not reset_at:
    ts = env.rollout(T, actions=actions)
    res = {k: getattr(ts, k).cpu().numpy() for k in out}
  else:
    for t in range(T):
      take(env.reset() if t in reset_at else env.step(actions[t]))
    res = {k: np.stack(v) for k, v in out.items()}
  res['info'] = {k: v.cpu().numpy() for k, v in env.bsuite_info().items()}
  env.close()
  return res


def _compare(meta, data, res):
  exact = meta['env_class'] not in cf.FLOAT_FAMILIES
  st_ref = data['step_type']
  np.testing.assert_array_equal(res['step_type'], st_ref)
  first = st_ref == 0
  # FIRST lanes: reference None -> engine 0 / 0
  assert np.all(res['reward'][first] == 0.0) and np.all(res['discount'][first] == 0.0)
  np.testing.assert_array_equal(res['discount'][~first], data['discount'][~first])
  r_ref, r_eng = data['reward'][~first], res['reward'][~first]
  if exact and meta['wrapper'] != 'noise' and not (meta['env_class'] == 'deep_sea' and not meta['kwargs'].get('deterministic', True)):
    np.testing.assert_array_equal(r_eng, r_ref)
  elif exact:
    # gaussian noise goes through log(): CUDA log vs glibc log may differ in the last ulp
    tol = 0.0 if res.get('host') else 1e-12
    np.testing.assert_allclose(r_eng, r_ref, rtol=tol, atol=tol)
  else:
    np.testing.assert_allclose(r_eng, r_ref, rtol=0, atol=cf.FLOAT_TOL * max(1.0, abs(meta['wrapper_arg']) if meta['wrapper'] == 'scale' else 1.0))
  obs_ref = data['observation'].reshape(res['observation'].shape)
  if exact:
    np.testing.assert_array_equal(res['observation'], obs_ref)
  else:
    np.testing.assert_allclose(res['observation'], obs_ref, rtol=0, atol=cf.FLOAT_TOL)
  for k, name in enumerate(meta['info_names']):
    ref = data['info'][:, k]
    if exact:
      np.testing.assert_array_equal(res['info'][name], ref, err_msg=name)
    else:
      np.testing.assert_allclose(res['info'][name], ref, rtol=0, atol=1e-6 * max(1.0, np.abs(ref).max()), err_msg=name)


def _needs_mnist(meta):
  return meta['env_class'] == 'mnist'


@pytest.mark.parametrize('fused', [False, True], ids=['stepwise', 'rollout'])
@pytest.mark.parametrize('name', cf.golden_case_names())
def test_host_path_matches_reference(name, fused, mnist_dir):
  meta, data = cf.load_golden(name)
  res = _run_engine(meta, data, 'cpu', fused)
  res['host'] = True
  _compare(meta, data, res)


@pytest.mark.gpu
@pytest.mark.parametrize('fused', [False, True], ids=['stepwise', 'rollout'])
@pytest.mark.parametrize('name', cf.golden_case_names())
def test_cuda_matches_reference(name, fused, mnist_dir):
  meta, data = cf.load_golden(name)
  res = _run_engine(meta, data, 'cuda', fused)
  _compare(meta, data, res)
