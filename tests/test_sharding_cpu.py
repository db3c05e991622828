"""CPU (gloo, world_size 2 and 3): the frame-sharding plan and the packed allreduce that combine
per-rank partial statistics -- the host-side logic of the multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_numpy as ora
from telluride_decoding_b200 import sharding


def test_shard_rows_partitions_every_row_once():
  keep = [7680, 5000, 12345, 64, 9000]
  first = [0, 100, 0, 0, 33]
  for world in (1, 2, 3, 4, 8):
    covered = [np.zeros(k, dtype=int) for k in keep]
    sizes = []
    for rank in range(world):
      f, k = sharding.shard_rows(first, keep, rank, world)
      n = 0
      for s, (a, b) in enumerate(zip(f, k)):
        assert first[s] <= a <= b <= keep[s]
        covered[s][a:b] += 1
        n += b - a
        # cuts inside a trial sit on 32-frame tile boundaries
        if a not in (first[s], keep[s]):
          assert a % 32 == 0
      sizes.append(n)
    for s in range(len(keep)):
      assert np.all(covered[s][first[s]:] == 1) and np.all(covered[s][:first[s]] == 0)
    total = sum(k - f for f, k in zip(first, keep))
    assert max(sizes) - min(sizes) <= 64 + total // (100 * world)      # balanced to a tile or so


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, result_path):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  rng = np.random.RandomState(0)                 # same data on every rank
  lens, pre, post, c = [700, 333, 1200], 2, 3, 4
  xs = [rng.randn(n, c).astype(np.float32) for n in lens]
  ys = [rng.randn(n, 1).astype(np.float32) for n in lens]
  keep = [700, 300, 1200]                        # the middle trial loses rows to drop_remainder
  first, mine = sharding.shard_rows(None, keep, rank, world)
  d = (pre + 1 + post) * c
  gram = torch.zeros((d, d), dtype=torch.float32)
  xtv = torch.zeros((d, 2), dtype=torch.float64)
  vtv = torch.zeros((2, 2), dtype=torch.float64)
  for x, y, a, b in zip(xs, ys, first, mine):
    rows = ora.add_context(x, pre, post)[a:b].astype(np.float64)     # halo comes from the same trial
    v = np.concatenate([y[a:b].astype(np.float64), np.ones((b - a, 1))], axis=1)
    gram += torch.from_numpy((rows.T @ rows).astype(np.float32))
    xtv += torch.from_numpy(rows.T @ v)
    vtv += torch.from_numpy(v.T @ v)
  assert sharding.is_distributed()
  sharding.allreduce_sum_([gram, xtv, vtv])
  if rank == 0:
    full = np.concatenate([ora.add_context(x, pre, post)[:k] for x, k in zip(xs, keep)]).astype(np.float64)
    vfull = np.concatenate([np.concatenate([y[:k].astype(np.float64), np.ones((k, 1))], axis=1)
                            for y, k in zip(ys, keep)])
    ok = (np.allclose(gram.numpy(), full.T @ full, rtol=1e-5, atol=1e-3) and
          np.allclose(xtv.numpy(), full.T @ vfull, rtol=1e-12) and
          np.allclose(vtv.numpy(), vfull.T @ vfull, rtol=1e-12) and vtv[1, 1].item() == sum(keep))
    with open(result_path, 'w') as f:
      f.write('ok' if ok else 'mismatch')
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_partial_statistics_allreduce_gloo(tmp_path, world):
  result = str(tmp_path / 'result.txt')
  mp.spawn(_worker, args=(world, _free_port(), result), nprocs=world, join=True)
  assert open(result).read() == 'ok'


def test_needed_row_range_covers_rows_and_halo():
  """Each rank uploads one contiguous range of raw rows: its design rows plus pre / post frames of
  the SAME trial on either side, clipped at the trial edges (zero padding there, brain_data.py:448-455)."""
  lens = [1000, 640, 2000]
  starts = [0, 1000, 1640]
  pre, post = 3, 17
  for world in (1, 2, 3, 8):
    for rank in range(world):
      first, keep = sharding.shard_rows(None, lens, rank, world)
      lo, hi = sharding.needed_row_range(starts, lens, lens, pre, post, rank, world)
      touched = [(st + max(f - pre, 0), st + min(k + post, n))
                 for st, n, f, k in zip(starts, lens, first, keep) if k > f]
      if not touched:
        assert (lo, hi) == (0, 0)
        continue
      assert lo == min(a for a, _ in touched) and hi == max(b for _, b in touched)
      assert 0 <= lo < hi <= sum(lens)
  assert sharding.needed_row_range(starts, lens, lens, pre, post, 0, 1) == (0, sum(lens))
