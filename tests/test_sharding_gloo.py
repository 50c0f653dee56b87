"""N > 1 path on CPU: world_size-2 gloo processes, each owning half of the lanes (host path of the C ABI).

Checks (i) sharding invariance -- rank r's lanes reproduce lanes [r*B/2, (r+1)*B/2) of an unsharded batch bit
for bit, because RNG keys depend on the global lane id only; (ii) the one collective of the path, the all-gather
of per-rank episode statistics, returns what the unsharded environment reports.
"""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BSUITE_ID, GLOBAL_BATCH, STEPS, SEED = 'catch_noise/7', 24, 60, 11


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from bsuite_b200 import distributed as bd
  env = bd.load_sharded(BSUITE_ID, GLOBAL_BATCH, device='cpu', seed=SEED, track_episodes=True, reward_dtype='float64')
  first, count = bd.shard_range(GLOBAL_BATCH, rank, world)
  assert (env.lane_offset, env.batch) == (first, count)
  actions = torch.as_tensor(np.random.RandomState(5).randint(3, size=(STEPS, GLOBAL_BATCH)).astype(np.int32))
  ts = env.rollout(STEPS, actions=actions[:, first:first + count].contiguous())
  gathered = bd.gather_episode_returns(env)
  rewards = bd.gather_lane_tensor(ts.reward.transpose(0, 1).contiguous())     # [B, T]
  np.savez(os.path.join(out_dir, f'rank{rank}.npz'), reward=ts.reward.numpy(), obs=ts.observation.numpy(),
           step_type=ts.step_type.numpy(), g_steps=gathered['steps'].numpy(), g_episode=gathered['episode'].numpy(),
           g_return=gathered['total_return'].numpy(), g_lanes=gathered['lanes'].numpy(),
           all_rewards=rewards.numpy())
  dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(tmp_path):
  import bsuite_b200
  from bsuite_b200 import distributed as bd
  world, port = 2, _free_port()
  mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  env = bsuite_b200.load_from_id(BSUITE_ID, batch=GLOBAL_BATCH, device='cpu', seed=SEED, track_episodes=True,
                                 reward_dtype='float64')
  actions = torch.as_tensor(np.random.RandomState(5).randint(3, size=(STEPS, GLOBAL_BATCH)).astype(np.int32))
  ts = env.rollout(STEPS, actions=actions)
  stats = env.episode_stats()
  ranks = [np.load(tmp_path / f'rank{r}.npz') for r in range(world)]
  for r, data in enumerate(ranks):
    first, count = bd.shard_range(GLOBAL_BATCH, r, world)
    np.testing.assert_array_equal(data['reward'], ts.reward.numpy()[:, first:first + count])
    np.testing.assert_array_equal(data['obs'], ts.observation.numpy()[:, first:first + count])
    np.testing.assert_array_equal(data['step_type'], ts.step_type.numpy()[:, first:first + count])
    # every rank holds the same gathered block
    np.testing.assert_array_equal(data['g_lanes'], [12.0, 12.0])
    np.testing.assert_array_equal(data['all_rewards'], ts.reward.numpy().T)
    for key, column in (('g_steps', 'steps'), ('g_episode', 'episode'), ('g_return', 'total_return')):
      want = [float(stats[column][bd.shard_range(GLOBAL_BATCH, q, world)[0]:][:12].sum()) for q in range(world)]
      np.testing.assert_allclose(data[key], want, rtol=0, atol=1e-12)


def test_shard_range_covers_batch():
  from bsuite_b200 import distributed as bd
  for batch, world in ((10, 3), (8, 8), (65536, 8), (7, 2)):
    spans = [bd.shard_range(batch, r, world) for r in range(world)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == batch
    assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
  with pytest.raises(ValueError):
    bd.shard_range(4, 4, 4)


SWEEP_IDS = ['catch/0', 'deep_sea/0', 'memory_len/3', 'bandit_noise/2']


def _sweep_worker(rank, world, port, out_dir):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from bsuite_b200 import suite
  batch = suite.SweepBatch(SWEEP_IDS, lanes=10, device='cpu', seed=SEED, rank=rank, world=world)
  batch.rollout(40, action_seed=9)
  gathered = batch.gather_returns()                    # [world, n_ids, 3]
  np.save(os.path.join(out_dir, f'sweep{rank}.npy'), gathered.numpy())
  dist.destroy_process_group()


def test_two_rank_sweep_batch_gathers_what_one_rank_computes(tmp_path):
  """BASELINE config #5 across ranks: every id's lanes are split over the ranks (global lane ids key the RNG and the
  on-device action stream), and the all-gather of the per-id Logging sums adds up to the single-rank sums."""
  from bsuite_b200 import suite
  world, port = 2, _free_port()
  mp.spawn(_sweep_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  whole = suite.SweepBatch(SWEEP_IDS, lanes=10, device='cpu', seed=SEED)
  whole.rollout(40, action_seed=9)
  want = whole.gather_returns().numpy()[0]              # [n_ids, 3]
  blocks = [np.load(tmp_path / f'sweep{r}.npy') for r in range(world)]
  np.testing.assert_array_equal(blocks[0], blocks[1])    # every rank holds the same gathered tensor
  assert blocks[0].shape == (world, len(SWEEP_IDS), 3)
  np.testing.assert_allclose(blocks[0].sum(axis=0), want, rtol=0, atol=1e-9)
  assert np.all(want[:, 2] > 0)                          # steps were taken for every id


def test_async_log_point_matches_the_synchronous_gather_on_host():
  """distributed.LogPoint (ticketed, slot-reusing) returns exactly what the synchronous reduction returns."""
  import torch
  import bsuite_b200
  from bsuite_b200 import distributed as bd
  envs = [bsuite_b200.load_from_id(i, batch=24, device='cpu', seed=3, track_episodes=True) for i in ('catch/0', 'bandit/0')]
  lp = bd.LogPoint(envs, slots=2)
  tickets = []
  for round_ in range(5):
    for env in envs:
      env.rollout(7, action_seed=round_)
    tickets.append((lp.issue(), torch.stack([env.episode_stat_sums() for env in envs])))
    ticket, want = tickets[-1]
    assert torch.equal(lp.result(ticket)[0], want)
  with pytest.raises(ValueError):
    lp.result(tickets[0][0])            # slot long since reused
