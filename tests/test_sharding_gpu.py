"""GPU parity of the time-sharded path: every rank's tcgen05 Gram over its own row ranges
(`seg_first`/`seg_keep` cut MID-TRIAL by sharding.shard_rows), summed over ranks, equals the oracle's
X'X and the unsharded kernel.  Reference semantics: the zero padding of brain_data.py:448-455 exists
only at file / chunk edges, never at a rank cut -- a rank reads the lag halo from the neighbouring
frames of the same trial."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ref_numpy as ora
from telluride_decoding_b200 import ops
from telluride_decoding_b200 import sharding
from tests import gpu_util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _trials(seed, lens, c, c2=None):
  rng = np.random.RandomState(seed)
  x = (rng.randn(sum(lens), c) + 0.15).astype(np.float32)
  x2 = rng.randn(sum(lens), c2).astype(np.float32) if c2 else None
  y = rng.randn(sum(lens), 1).astype(np.float32)
  starts = np.concatenate(([0], np.cumsum(lens)[:-1])).tolist()
  return x, x2, y, starts


def _oracle_design(x, starts, lens, keep, pre, post):
  """float64 lagged rows [0, keep) of every trial, built by the oracle's add_context per trial."""
  parts = [ora.add_context(x[s:s + n], pre, post)[:k].astype(np.float64)
           for s, n, k in zip(starts, lens, keep)]
  return np.concatenate(parts, axis=0)


@pytest.mark.parametrize('world', [2, 3, 8])
@pytest.mark.parametrize('cta_group', [1, 2])
def test_sharded_gram_sums_to_the_oracle_gram(world, cta_group):
  """pre > 0 and post > 0, ragged kept rows, cuts in the middle of trials."""
  lens = [5000, 3333, 7001, 1200]
  keep = [5000, 3300, 6900, 1200]
  c, pre, post = 24, 5, 17                                   # D = 552
  x, _, y, starts = _trials(31 + world, lens, c)
  xd, yd = torch.as_tensor(x, device='cuda'), torch.as_tensor(y, device='cuda')
  sx = ops.Stream(xd, pre, post, starts, lens)
  sy = ops.Stream(yd, 0, 0, starts, lens)
  opt = ops.gram_options(cta_group=cta_group)
  whole = ops.lagged_gram([sx], keep, options=opt).double()
  t_whole, v_whole = ops.lagged_thin(sx, sy, keep)
  total = torch.zeros_like(whole)
  t_sum, v_sum = torch.zeros_like(t_whole), torch.zeros_like(v_whole)
  mid_trial_cuts = 0
  for rank in range(world):
    first, mine = sharding.shard_rows(None, keep, rank, world)
    mid_trial_cuts += sum(1 for f, k in zip(first, keep) if 0 < f < k)
    total += ops.lagged_gram([sx], mine, options=opt, seg_first=first).double()
    t, v = ops.lagged_thin(sx, sy, mine, seg_first=first)
    t_sum += t
    v_sum += v
  assert mid_trial_cuts >= world - 1 - (len(lens) - 1)       # the plan really cuts inside trials
  design = _oracle_design(x, starts, lens, keep, pre, post)
  want = design.T @ design
  assert gpu_util.rel_fro(total.cpu().numpy(), want) < 1e-5
  assert gpu_util.rel_fro(total.cpu().numpy(), whole.cpu().numpy()) < 5e-6
  ycat = np.concatenate([y[s:s + k] for s, k in zip(starts, keep)]).astype(np.float64)
  v = np.concatenate([ycat, np.ones_like(ycat)], axis=1)
  np.testing.assert_allclose(t_sum.cpu().numpy(), design.T @ v, rtol=1e-5, atol=1e-4)
  np.testing.assert_allclose(v_sum.cpu().numpy(), v.T @ v, rtol=1e-6)
  np.testing.assert_allclose(t_sum.cpu().numpy(), t_whole.cpu().numpy(), rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize('world', [2, 3, 8])
def test_sharded_two_stream_gram(world):
  """Z = [X | X2] with different contexts per stream on the CTA-pair kernel, rows cut mid-trial."""
  lens = [4100, 2900, 6000]
  keep = [4100, 2800, 5990]
  x, x2, _, starts = _trials(77, lens, 40, 6)
  xd, x2d = torch.as_tensor(x, device='cuda'), torch.as_tensor(x2, device='cuda')
  sx = ops.Stream(xd, 3, 22, starts, lens)                    # 26 lags x 40 = 1040 columns
  s2 = ops.Stream(x2d, 2, 5, starts, lens)                    # 8 lags x 6 = 48 columns
  whole = ops.lagged_gram([sx, s2], keep).double()
  total = torch.zeros_like(whole)
  for rank in range(world):
    first, mine = sharding.shard_rows(None, keep, rank, world)
    total += ops.lagged_gram([sx, s2], mine, seg_first=first).double()
  z = np.concatenate([_oracle_design(x, starts, lens, keep, 3, 22),
                      _oracle_design(x2, starts, lens, keep, 2, 5)], axis=1)
  assert gpu_util.rel_fro(total.cpu().numpy(), z.T @ z) < 1e-5
  assert gpu_util.rel_fro(total.cpu().numpy(), whole.cpu().numpy()) < 5e-6


def test_shard_with_nonzero_first_rows_and_unaligned_cut():
  """A shard of a shard: seg_first > 0 to begin with, cuts NOT on k-tile boundaries (align=1)."""
  lens = [3000, 2000]
  first0, keep = [37, 0], [2950, 1999]
  x, _, _, starts = _trials(5, lens, 16)
  sx = ops.Stream(torch.as_tensor(x, device='cuda'), 2, 9, starts, lens)
  total = None
  for rank in range(3):
    first, mine = sharding.shard_rows(first0, keep, rank, 3, align=1)
    g = ops.lagged_gram([sx], mine, seg_first=first, options=ops.gram_options(cta_group=2)).double()
    total = g if total is None else total + g
  parts = [ora.add_context(x[s:s + n], 2, 9)[f:k].astype(np.float64)
           for s, n, f, k in zip(starts, lens, first0, keep)]
  design = np.concatenate(parts, axis=0)
  assert gpu_util.rel_fro(total.cpu().numpy(), design.T @ design) < 1e-5


def test_two_process_nccl_sharded_statistics():
  """Two ranks over NCCL (tools/dist_check.py under torchrun): sharded statistics + allreduce equal
  the single-GPU statistics and weights.  Skips on a box with one GPU."""
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
         '--master-addr', '127.0.0.1', '--master-port', '29731',
         os.path.join(ROOT, 'tools', 'dist_check.py')]
  res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
  assert res.stdout.count('OK') == 2


def test_packed_triangle_roundtrip_and_single_rank_comm():
  """td_pack_upper_f32 / td_unpack_upper_f32 (what td_allreduce_sym_f32 sends: n(n+1)/2 floats) and
  the communicator life cycle with one rank (NCCL bound by libtdgram.so at run time)."""
  import ctypes
  from telluride_decoding_b200 import _lib
  lib = _lib.load()
  n = 1003
  a = torch.randn((n, n), device='cuda')
  g = (a + a.t()).contiguous()
  packed = torch.empty(n * (n + 1) // 2, device='cuda')
  st = _lib.current_stream_ptr()
  _lib.check(lib.td_pack_upper_f32(_lib.ptr(g), ctypes.c_int64(n), ctypes.c_int64(n), _lib.ptr(packed), st))
  rows, cols = torch.triu_indices(n, n, device='cuda')
  assert torch.equal(packed, g[rows, cols])
  out = torch.zeros_like(g)
  _lib.check(lib.td_unpack_upper_f32(_lib.ptr(packed), ctypes.c_int64(n), ctypes.c_int64(n), _lib.ptr(out), st))
  assert torch.equal(out, g)
  ident = (ctypes.c_ubyte * 128)()
  _lib.check(lib.td_comm_unique_id(ident))
  comm = ctypes.c_void_p()
  _lib.check(lib.td_comm_init(ctypes.byref(comm), ident, 0, 1))
  version = ctypes.c_int(0)
  _lib.check(lib.td_comm_nccl_version(ctypes.byref(version)))
  assert version.value >= 20000
  before = g.clone()
  _lib.check(lib.td_allreduce_sym_f32(comm, _lib.ptr(g), ctypes.c_int64(n), ctypes.c_int64(n), _lib.ptr(packed),
                                      ctypes.c_size_t(packed.numel() * 4), st))
  _lib.check(lib.td_allreduce_sum_f32(comm, _lib.ptr(g), ctypes.c_int64(g.numel()), st))
  torch.cuda.synchronize()
  assert torch.equal(g, before)                       # one rank: the sum is the input
  _lib.check(lib.td_comm_destroy(comm))
  # too small a packed buffer is reported, not overrun (needs a multi-rank comm to get that far: world 1 returns early)
  assert lib.td_pack_upper_f32(None, ctypes.c_int64(4), ctypes.c_int64(4), None, st) != 0


def test_sharded_upload_of_a_rank_local_host_array():
  """TrialBrainData.set_trials(upload_rows=...) with pinned host arrays that hold ONLY the rows a rank
  touches (its design rows + lag halo, sharding.needed_row_range): the device arrays carry those rows
  at their global positions, zeros elsewhere, and the rank's partial Gram equals the one computed from
  the whole recording."""
  from telluride_decoding_b200 import brain_data
  lens = [3000, 2500, 4000]
  c, pre, post = 16, 3, 12
  x, _, y, starts = _trials(3, lens, c)
  names = ['a', 'b', 'c']
  world = 3
  whole = None
  for rank in range(world):
    lo, hi = sharding.needed_row_range(starts, lens, lens, pre, post, rank, world)
    xp = torch.from_numpy(x[lo:hi].copy()).pin_memory()
    yp = torch.from_numpy(y[lo:hi].copy()).pin_memory()
    bd = brain_data.TrialBrainData('eeg', 'env', 100.0, pre_context=pre, post_context=post,
                                   final_batch_size=500, shuffle_buffer_size=0)
    bd.set_trials(names, lens, xp, yp, upload_rows=(lo, hi))
    ds = bd.create_dataset('train')
    sx = ds.stream('x')
    torch.cuda.synchronize()
    dev = sx.data.cpu().numpy()
    np.testing.assert_array_equal(dev[lo:hi], x[lo:hi])
    assert not dev[:lo].any() and not dev[hi:].any()
    order = sorted(range(len(sx.seg_start)), key=lambda i: sx.seg_start[i])
    f_sorted, k_sorted = sharding.shard_rows(None, [ds.seg_rows[i] for i in order], rank, world)
    first, keep = [0] * 3, [0] * 3
    for pos, i in enumerate(order):
      first[i], keep[i] = f_sorted[pos], k_sorted[pos]
    g = ops.lagged_gram([sx], keep, seg_first=first).double()
    whole = g if whole is None else whole + g
  design = _oracle_design(x, starts, lens, [n - n % 1 for n in lens], pre, post)
  # drop_remainder keeps 9500 of 9500 rows (500 | 9500): all rows
  assert gpu_util.rel_fro(whole.cpu().numpy(), design.T @ design) < 1e-5
