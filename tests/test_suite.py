"""SweepBatch (BASELINE config #5): the heterogeneous 23-experiment batch, checked per id against the oracle."""

import numpy as np
import pytest

import bsuite_b200
from bsuite_b200 import suite
from oracle import bsuite_oracle as oracle

DEVICES = [pytest.param('cpu', id='host'), pytest.param('cuda', id='cuda', marks=pytest.mark.gpu)]

# experiment -> (environment class, kwargs, wrapper, wrapper_arg) for setting 0
_SETTING0 = {
    'bandit/0': ('bandit', dict(mapping_seed=0), None, 0.),
    'bandit_noise/0': ('bandit', dict(mapping_seed=0), 'noise', 0.1),
    'bandit_scale/0': ('bandit', dict(mapping_seed=0), 'scale', 0.001),
    'catch/0': ('catch', {}, None, 0.),
    'catch_scale/0': ('catch', {}, 'scale', 0.001),
    'deep_sea/0': ('deep_sea', dict(size=10, mapping_seed=42), None, 0.),
    'discounting_chain/0': ('discounting_chain', dict(mapping_seed=0), None, 0.),
    'umbrella_length/0': ('umbrella_chain', dict(chain_length=1, n_distractor=20), None, 0.),
}


def test_one_per_experiment_covers_all_23():
  ids = suite.one_per_experiment()
  assert len(ids) == 23 and len({i.split('/')[0] for i in ids}) == 23


@pytest.mark.parametrize('device', DEVICES)
def test_sweep_batch_matches_oracle_per_id(device, mnist_dir):
  T, lanes, seed = 25, 40, 3
  batch = suite.SweepBatch(lanes=lanes, device=device, seed=seed)
  assert len(batch.envs) == 23
  result = batch.rollout(T, action_seed=7)
  returns = batch.gather_returns()
  assert tuple(returns.shape) == (1, 23, 3)
  for k, bsuite_id in enumerate(batch.bsuite_ids):
    ts = result[bsuite_id]
    env = batch.envs[bsuite_id]
    assert tuple(ts.observation.shape) == (T, lanes) + tuple(env.obs_shape)
    st = ts.step_type.cpu().numpy()
    assert np.all(st[0] == 0)
    np.testing.assert_allclose(float(returns[0, k, 2]), float((st != 0).sum()))
    np.testing.assert_allclose(float(returns[0, k, 1]), float((st == 2).sum()))
    if bsuite_id in _SETTING0:
      env_class, kwargs, wrapper, arg = _SETTING0[bsuite_id]
      actions = batch.last_buffers(bsuite_id).actions.cpu().numpy()   # pylint: disable=protected-access
      want = oracle.run_lanes(env_class, kwargs, actions[:, :6], seed=seed, wrapper=wrapper, wrapper_arg=arg)
      np.testing.assert_array_equal(st[:, :6], want['step_type'])
      np.testing.assert_array_equal(ts.observation.cpu().numpy()[:, :6], want['observation'])
      if wrapper != 'noise':
        np.testing.assert_array_equal(ts.reward.cpu().numpy()[:, :6], want['reward'].astype(np.float32))
  batch.close()


def test_sweep_batch_shards_every_id_evenly():
  a = suite.SweepBatch(bsuite_ids=['catch/0', 'bandit/3'], lanes=10, device='cpu', seed=1, rank=1, world=4)
  assert (a.local_lanes, a.lane_offset) == (3, 3) and all(e.batch == 3 and e.lane_offset == 3 for e in a.envs.values())
