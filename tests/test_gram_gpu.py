"""GPU parity: the tcgen05 lagged Gram (both tile modes), the thin products, the forward and the
row materialiser against the oracle; size-independent properties at larger shapes."""
import numpy as np
import pytest
import torch

from oracle import ref_numpy as ora
from telluride_decoding_b200 import ops
from tests import gpu_util

pytestmark = pytest.mark.gpu

GRAM_TOL = 1e-5     # BASELINE.json north_star: relative Frobenius error of the Gram


def _dataset(n, c, pre, post, batch, seed, c2=1, pre2=0, post2=0, offset=0, chunk=1000000, dy=1):
  rng = np.random.RandomState(seed)
  x = (rng.randn(n, c) + 0.2).astype(np.float32)
  x2 = rng.randn(n, c2).astype(np.float32)
  y = rng.randn(n, dy).astype(np.float32)
  bd = gpu_util.make_test_data(x, y, x2, pre_context=pre, post_context=post, in2_pre_context=pre2,
                               in2_post_context=post2, input_offset=offset,
                               initial_batch_size=chunk, final_batch_size=batch)
  ds = bd.create_dataset('train')
  batches = gpu_util.oracle_batches(x, y, x2, None, pre, post, pre2, post2, offset, batch, chunk)
  return ds, batches


# (cta_group, method): 128x128 tiles; CTA pairs with the lagged operand built on chip from a raw tile +
# lag halo (method 4, the default for pairs); CTA pairs with staged planes and the A operand in tensor
# memory (3); CTA pairs with both operands in shared memory (2).
@pytest.mark.parametrize('cta_group,method', [(1, 0), (2, 0), (2, 3), (2, 2)])
@pytest.mark.parametrize('n,c,pre,post,batch', [
    (1000, 2, 1, 1, 100),       # C=2, L=3 (SURVEY build-plan shape)
    (1000, 3, 1, 2, 100),       # golden ridge_small shape
    (3000, 32, 16, 16, 500),    # C=32, L=33 (test/brain_model_test.py:728-776)
    (7680, 64, 0, 31, 768),     # C=64, L=32 (config 2 shape, one trial)
    (2501, 8, 0, 7, 300),       # rows lost to drop_remainder (tail mask)
    (4100, 148, 3, 3, 1000),    # MEG fixture width, channel padding 148 -> 160
])
def test_gram_matches_oracle(n, c, pre, post, batch, cta_group, method):
  ds, batches = _dataset(n, c, pre, post, batch, seed=n + c)
  (w, keep), = ds.passes()
  g = ops.lagged_gram([ds.stream('x')], keep,
                      options=ops.gram_options(method=method, cta_group=cta_group)).cpu().numpy()
  x64 = gpu_util.design64(batches)
  want = x64.T @ x64
  assert gpu_util.rel_fro(g, want) < GRAM_TOL
  np.testing.assert_array_equal(g, g.T)            # exactly symmetric, like numpy's syrk path
  # the reference's own float32 accumulation is no closer to float64 than we are x 30
  ref32 = 0
  for b in batches:
    xb = np.asarray(b[0]['input_1'])
    ref32 = ref32 + np.matmul(xb.T, xb)
  assert gpu_util.rel_fro(g, ref32) < GRAM_TOL


@pytest.mark.parametrize('method', [0, 3, 2])
@pytest.mark.parametrize('chunk_rows,superchunk_rows', [(32, 4096), (64, 1024), (96, 2048), (128, 512)])
def test_pair_kernels_with_short_chunks_and_many_work_items(method, chunk_rows, superchunk_rows):
  """Schedule stress for the CTA-pair kernels: one-k-tile flush periods and tiny superchunks make
  every accumulator hand-off and work-item boundary as frequent as possible (with 32-frame chunks
  one half of the tile can run two chunks ahead of the other; a hand-off barrier shared by both
  halves deadlocked there)."""
  rng = np.random.RandomState(17)
  n_trials, frames, c, lags = 5, 3111, 40, 32           # D = 1280: 5 x 5 pair tiles, ragged trial ends
  x = (rng.randn(n_trials * frames, c) + 0.1).astype(np.float32)
  xd = torch.as_tensor(x, device='cuda')
  starts = [i * frames for i in range(n_trials)]
  lens = [frames] * n_trials
  keep = [frames - 7 * i for i in range(n_trials)]
  stream = ops.Stream(xd, 0, lags - 1, starts, lens)
  want = ops.lagged_gram([stream], keep, options=ops.gram_options(method=1))
  for _ in range(3):
    got = ops.lagged_gram([stream], keep, options=ops.gram_options(
        method=method, cta_group=2, chunk_rows=chunk_rows, superchunk_rows=superchunk_rows))
    assert gpu_util.rel_fro(got.cpu().numpy(), want.cpu().numpy()) < GRAM_TOL


@pytest.mark.parametrize('method', [0, 3, 2])
def test_two_stream_gram_on_the_pair_kernels(method):
  """Z = [X | X2] wide enough for CTA pairs (two tensor maps, channel padding 40 -> 64 and 6 -> 8
  boxes, different contexts per stream, ragged kept rows) against the fp32 cross-check kernel."""
  rng = np.random.RandomState(23)
  n_trials, frames = 4, 2500
  x = (rng.randn(n_trials * frames, 40) + 0.2).astype(np.float32)
  x2 = rng.randn(n_trials * frames, 6).astype(np.float32)
  xd, x2d = torch.as_tensor(x, device='cuda'), torch.as_tensor(x2, device='cuda')
  starts = [i * frames for i in range(n_trials)]
  lens = [frames] * n_trials
  keep = [frames, frames - 100, frames, frames - 1]
  sx = ops.Stream(xd, 3, 22, starts, lens)            # 26 lags x 40 = 1040 columns
  s2 = ops.Stream(x2d, 0, 5, starts, lens)            # 6 lags x 6 = 36 columns
  want = ops.lagged_gram([sx, s2], keep, options=ops.gram_options(method=1)).cpu().numpy()
  got = ops.lagged_gram([sx, s2], keep, options=ops.gram_options(method=method, cta_group=2)).cpu().numpy()
  assert got.shape == (1076, 1076)
  assert gpu_util.rel_fro(got, want) < GRAM_TOL
  np.testing.assert_array_equal(got, got.T)


def test_simt_cross_check_agrees():
  ds, batches = _dataset(2000, 20, 2, 2, 200, seed=5)
  keep = ds.seg_rows
  g0 = ops.lagged_gram([ds.stream('x')], keep).cpu().numpy()
  g1 = ops.lagged_gram([ds.stream('x')], keep, options=ops.gram_options(method=1)).cpu().numpy()
  x64 = gpu_util.design64(batches)
  assert gpu_util.rel_fro(g1, x64.T @ x64) < 1e-6
  assert gpu_util.rel_fro(g0, g1) < GRAM_TOL


@pytest.mark.parametrize('offset,chunk', [(0, 1000000), (3, 1000000), (-2, 1000000), (0, 700), (2, 640)])
def test_two_stream_gram_offsets_and_chunks(offset, chunk):
  """Z = [X | Y] with different contexts, input_offset and chunk-edge zero padding
  (brain_data.py:466-483, :492)."""
  ds, batches = _dataset(2600, 12, 1, 3, 100, seed=9, c2=5, pre2=2, post2=1, offset=offset,
                         chunk=chunk)
  (w, keep), = ds.passes()
  g = ops.lagged_gram([ds.stream('x'), ds.stream('x2')], keep).cpu().numpy()
  z = np.concatenate([gpu_util.design64(batches), gpu_util.design64(batches, 'input_2')], axis=1)
  assert g.shape == (z.shape[1], z.shape[1])
  assert gpu_util.rel_fro(g, z.T @ z) < GRAM_TOL


def test_thin_products_and_rows_and_predict():
  ds, batches = _dataset(1500, 7, 2, 1, 250, seed=3, dy=3)
  (w, keep), = ds.passes()
  xtv, vtv = ops.lagged_thin(ds.stream('x'), ds.stream('y'), keep)
  x64 = gpu_util.design64(batches)
  y64 = np.concatenate([np.asarray(b[1], dtype=np.float64) for b in batches], axis=0)
  v = np.concatenate([y64, np.ones((y64.shape[0], 1))], axis=1)
  # fp32 products carried into float64 every 32 frames
  np.testing.assert_allclose(xtv.cpu().numpy(), x64.T @ v, rtol=2e-6, atol=2e-4)
  np.testing.assert_allclose(vtv.cpu().numpy(), v.T @ v, rtol=1e-10, atol=1e-9)
  assert vtv[-1, -1].item() == 1500
  # the lagged rows are bit-identical to the reference's tf.signal.frame layout
  got = [b for b in ds]
  assert len(got) == len(batches) == 6
  for (gd, gy), (od, oy) in zip(got, batches):
    np.testing.assert_array_equal(gd['input_1'].cpu().numpy(), od['input_1'])
    np.testing.assert_array_equal(gd['input_2'].cpu().numpy(), od['input_2'])
    np.testing.assert_array_equal(gy.cpu().numpy(), oy)
  # forward kernel = Dense on the lagged rows
  rng = np.random.RandomState(1)
  wmat = rng.randn(x64.shape[1], 3).astype(np.float32)
  bias = rng.randn(3).astype(np.float32)
  mean = rng.randn(x64.shape[1]).astype(np.float32)
  out = ops.lagged_predict(ds.stream('x'), ds.seg_rows, ds.seg_out_rows(),
                           torch.as_tensor(wmat, device='cuda'), torch.as_tensor(bias, device='cuda'),
                           torch.as_tensor(mean, device='cuda'), total_rows=ds.total_rows)
  want = (x64 - mean) @ wmat + bias
  np.testing.assert_allclose(out.cpu().numpy(), want, rtol=1e-4, atol=1e-4)


def test_lag_goldens_through_the_kernel():
  """test/brain_data_test.py:291-294, :317-320 via td_lagged_rows_f32."""
  t = np.arange(40, dtype=np.float32)
  x = np.stack([t, 1000 + t], axis=1)
  y = (3000 + t).reshape(-1, 1)
  for pre, post, want in ((2, 0, [[0, 0, 0, 0, 0, 1000], [0, 0, 0, 1000, 1, 1001],
                                  [0, 1000, 1, 1001, 2, 1002]]),
                          (0, 2, [[0, 1000, 1, 1001, 2, 1002], [1, 1001, 2, 1002, 3, 1003],
                                  [2, 1002, 3, 1003, 4, 1004]])):
    bd = gpu_util.make_test_data(x, y, pre_context=pre, post_context=post, final_batch_size=20)
    (inputs, out), = bd.create_dataset('program_test').take(1)
    np.testing.assert_array_equal(inputs['input_1'][:3].cpu().numpy(), want)
    np.testing.assert_array_equal(out[:3, 0].cpu().numpy(), [3000, 3001, 3002])
    assert bd.input_fields_width() == 6


def test_segment_additivity_and_toeplitz_structure_at_scale():
  """Properties that hold at any size: G(all trials) = sum of per-trial G, and every interior
  lag block equals the channel autocorrelation at that lag."""
  rng = np.random.RandomState(11)
  c, post, n_trials, n = 64, 31, 6, 7680
  x = rng.randn(n_trials * n, c).astype(np.float32)
  xd = torch.as_tensor(x, device='cuda')
  starts = [i * n for i in range(n_trials)]
  lens = [n] * n_trials
  full = ops.lagged_gram([ops.Stream(xd, 0, post, starts, lens)], lens).to(torch.float64)
  parts = 0
  for s in starts:
    parts = parts + ops.lagged_gram([ops.Stream(xd, 0, post, [s], [n])], [n]).to(torch.float64)
  assert gpu_util.rel_fro(full.cpu().numpy(), parts.cpu().numpy()) < 5e-6
  # block (k, k') with k' - k = tau: sum over trials of x[u]' x[u + tau] for rows with both in range
  x64 = x.astype(np.float64).reshape(n_trials, n, c)
  k, tau = 5, 9
  want = sum(x64[i, k:n - tau].T @ x64[i, k + tau:n] for i in range(n_trials))
  blk = full[k * c:(k + 1) * c, (k + tau) * c:(k + tau + 1) * c].cpu().numpy()
  assert gpu_util.rel_fro(blk, want) < 1e-4      # off-diagonal block: sqrt(T)-sized entries
  diag = full[k * c:(k + 1) * c, k * c:(k + 1) * c].cpu().numpy()
  want_d = sum(x64[i, k:n].T @ x64[i, k:n] for i in range(n_trials))
  assert gpu_util.rel_fro(diag, want_d) < GRAM_TOL


def _lag_blocks_match(G, xd, n_trials, n, c, lags, pairs, tol):
  """Lag block (k1, k2) of the Gram of a pre=0 design = sum over trials of x[k1:k1+h]' x[k2:k2+h],
  h = n - max(k1, k2) (float64 on the device as the independent check)."""
  x3 = xd.view(n_trials, n, c)
  for k1, k2 in pairs:
    h = n - max(k1, k2)
    want = torch.zeros((c, c), dtype=torch.float64, device=xd.device)
    for i in range(n_trials):
      want += x3[i, k1:k1 + h].to(torch.float64).t() @ x3[i, k2:k2 + h].to(torch.float64)
    got = G[k1 * c:(k1 + 1) * c, k2 * c:(k2 + 1) * c].to(torch.float64)
    err = float(torch.linalg.norm(got - want) / torch.linalg.norm(want))
    assert err < tol, (k1, k2, err)


def test_full_size_config2_recording():
  """BASELINE.json configs[1] at full size (60 x 7680 frames, 64 channels, 32 lags, D = 2048):
  exact symmetry, additivity over a split of the trials, and lag blocks against float64."""
  gen = torch.Generator(device='cuda')
  gen.manual_seed(2)
  n_trials, n, c, lags = 60, 7680, 64, 32
  xd = torch.randn((n_trials * n, c), generator=gen, device='cuda')
  starts = [i * n for i in range(n_trials)]
  lens = [n] * n_trials
  G = ops.lagged_gram([ops.Stream(xd, 0, lags - 1, starts, lens)], lens)
  assert torch.equal(G, G.t())
  _lag_blocks_match(G, xd, n_trials, n, c, lags, [(0, 0), (3, 20), (31, 31), (0, 31), (17, 18)], 2e-5)
  keep_a = [n if i < 23 else 0 for i in range(n_trials)]
  keep_b = [0 if i < 23 else n for i in range(n_trials)]
  Ga = ops.lagged_gram([ops.Stream(xd, 0, lags - 1, starts, lens)], keep_a).to(torch.float64)
  Gb = ops.lagged_gram([ops.Stream(xd, 0, lags - 1, starts, lens)], keep_b).to(torch.float64)
  assert gpu_util.rel_fro((Ga + Gb).cpu().numpy(), G.to(torch.float64).cpu().numpy()) < 5e-6


def test_config4_width_one_trial():
  """BASELINE.json configs[3] width (306 channels x 100 lags, D = 30 600, 7875 pair tiles) on one
  100 000-frame trial: thousands of work items per CTA pair, a boundary every 156 k-tiles -- the
  shape on which a shared accumulator hand-off barrier used to deadlock."""
  free, _ = torch.cuda.mem_get_info()
  if free < 12 * (1 << 30):
    pytest.skip('needs 12 GB of device memory')
  gen = torch.Generator(device='cuda')
  gen.manual_seed(4)
  n, c, lags = 100000, 306, 100
  xd = torch.randn((n, c), generator=gen, device='cuda')
  G = ops.lagged_gram([ops.Stream(xd, 0, lags - 1, [0], [n])], [n])
  _lag_blocks_match(G, xd, 1, n, c, lags, [(0, 0), (3, 50), (99, 99), (17, 18), (0, 99)], 2e-5)
  assert torch.equal(G[:4096, :4096], G[:4096, :4096].t())
  del G
  torch.cuda.empty_cache()


def test_argument_errors_surface_as_exceptions():
  x = torch.zeros((100, 4), device='cuda')
  from telluride_decoding_b200 import _lib
  with pytest.raises(_lib.TdError, match='n_streams'):
    s = ops.Stream(x, 0, 0, [0], [100])
    ops.lagged_gram([s, s, s], [100])
  with pytest.raises(ValueError, match='weights have'):
    ops.lagged_predict(ops.Stream(x, 0, 1, [0], [100]), [100], [0], torch.zeros((3, 1), device='cuda'))


@pytest.mark.parametrize('case', [
    dict(lens=[2000, 1777], c=10, pre=3, post=56),                  # row stride 40 B: TMA cannot address it -> padded copy
    dict(lens=[1500], c=306, pre=0, post=9),                        # cfg4 channel count (ld = 306)
    dict(lens=[2200, 900], c=60, pre=2, post=9, ld_pad=4),          # strided view of a wider array, zero-copy TMA
    dict(lens=[2048, 2048], c=32, pre=1, post=17),                  # 19 lags: last lag group has one real lag
    dict(lens=[700, 33, 5, 1200], c=16, pre=40, post=50),           # trials shorter than the lag span; masks over several k-tiles
    dict(lens=[3000, 2500], c=40, pre=5, post=26, keep=[2900, 2500], first=[37, 1024]),
    dict(lens=[64] * 200, c=24, pre=4, post=19),                    # hundreds of tiny trials
])
@pytest.mark.parametrize('method', [4, 5, 6])
def test_operand_built_on_chip_masks_and_layouts(case, method):
  """Method 4 specifics against the ORACLE's lagged design: zero padding at trial edges rebuilt by
  masks (the raw TMA box runs into the neighbouring trial; brain_data.py:448-452), row ranges,
  channel / lag padding of the 4-lag x 8-channel row groups, unaligned and strided inputs.
  Method 5 (block-Toeplitz structure: only the first block row is contracted, the rest follows along
  the lag diagonals with edge sums at the head and tail of every trial) on the same cases."""
  lens, c, pre, post = case['lens'], case['c'], case['pre'], case['post']
  keep = case.get('keep', lens)
  first = case.get('first')
  rng = np.random.RandomState(c + len(lens))
  x = (rng.randn(sum(lens), c + case.get('ld_pad', 0)) + 0.1).astype(np.float32)
  xd = torch.as_tensor(x, device='cuda')[:, :c]
  starts = np.concatenate(([0], np.cumsum(lens)[:-1])).tolist()
  stream = ops.Stream(xd, pre, post, starts, lens)
  got = ops.lagged_gram([stream], keep, options=ops.gram_options(method=method), seg_first=first)
  parts = []
  for s, n, k, f in zip(starts, lens, keep, first or [0] * len(lens)):
    parts.append(ora.add_context(x[s:s + n, :c], pre, post)[f:k].astype(np.float64))
  design = np.concatenate(parts, axis=0)
  assert gpu_util.rel_fro(got.cpu().numpy(), design.T @ design) < GRAM_TOL
  assert torch.equal(got, got.t())


def test_operand_built_on_chip_two_streams_with_offset():
  """Z = [X | X2] through method 4 with input_offset (the streams start at different raw rows and
  have different lengths per segment, brain_data.py:466-475) and chunk edges."""
  for offset, chunk in ((3, 1000000), (-2, 700)):
    ds, batches = _dataset(2600, 40, 1, 14, 100, seed=9, c2=5, pre2=2, post2=1, offset=offset, chunk=chunk)
    (w, keep), = ds.passes()
    z = np.concatenate([gpu_util.design64(batches), gpu_util.design64(batches, 'input_2')], axis=1)
    for method in (4, 6):
      g = ops.lagged_gram([ds.stream('x'), ds.stream('x2')], keep,
                          options=ops.gram_options(method=method)).cpu().numpy()
      assert gpu_util.rel_fro(g, z.T @ z) < GRAM_TOL, method


def test_block_toeplitz_gram_on_config2_and_sharded_rows():
  """Method 5 at the cfg 2 shape (one trial) against the dense kernel, and summed over the row ranges
  of three ranks (mid-trial cuts: the head / tail edge sums then sit in the middle of a trial)."""
  from telluride_decoding_b200 import sharding
  gen = torch.Generator(device='cuda')
  gen.manual_seed(12)
  n_trials, n, c, lags = 4, 7680, 64, 32
  xd = torch.randn((n_trials * n, c), generator=gen, device='cuda')
  starts = [i * n for i in range(n_trials)]
  lens = [n] * n_trials
  keep = [n, n - 128, n, n - 5]
  s = ops.Stream(xd, 3, lags - 4, starts, lens)
  dense = ops.lagged_gram([s], keep).double()
  fast = ops.lagged_gram([s], keep, options=ops.gram_options(method=5))
  assert torch.equal(fast, fast.t())
  assert gpu_util.rel_fro(fast.cpu().numpy(), dense.cpu().numpy()) < 5e-6
  total = torch.zeros_like(dense)
  for rank in range(3):
    first, mine = sharding.shard_rows(None, keep, rank, 3)
    total += ops.lagged_gram([s], mine, seg_first=first, options=ops.gram_options(method=5)).double()
  assert gpu_util.rel_fro(total.cpu().numpy(), dense.cpu().numpy()) < 5e-6
  with pytest.raises(Exception, match='one stream'):
    ops.lagged_gram([s, s], keep, options=ops.gram_options(method=5))


def test_per_segment_grams_dense_and_block_toeplitz():
  """One Gram per trial from one launch (leave-one-trial-out statistics): dense kernel and method 5
  against per-trial calls of the dense kernel."""
  gen = torch.Generator(device='cuda')
  gen.manual_seed(3)
  n_trials, n, c, lags = 7, 2048, 24, 12
  xd = torch.randn((n_trials * n, c), generator=gen, device='cuda')
  starts = [i * n for i in range(n_trials)]
  lens = [n] * n_trials
  s = ops.Stream(xd, 2, lags - 3, starts, lens)
  want = torch.stack([ops.lagged_gram([ops.Stream(xd, 2, lags - 3, [starts[i]], [n])], [n]) for i in range(n_trials)])
  for method in (0, 5, 6):
    got = ops.lagged_gram_per_segment([s], lens, options=ops.gram_options(method=method))
    assert got.shape == (n_trials, c * lags, c * lags)
    for i in range(n_trials):
      assert gpu_util.rel_fro(got[i].cpu().numpy(), want[i].cpu().numpy()) < 5e-6, (method, i)
      assert torch.equal(got[i], got[i].t())


def test_fp16_split_flavour_over_badly_scaled_channels():
  """Method 6 multiplies every channel by a power of two before the fp16 hi / lo split and undoes it in
  the epilogue: channels in volts next to channels in the thousands, a heavy-tailed channel (one
  artefact 1e4 times the rest), a constant and an all-zero channel -- block by block against float64
  products, each block normalised by its own size (a global Frobenius norm would only see the largest
  channels)."""
  rng = np.random.RandomState(5)
  n, c, pre, post = 6000, 24, 2, 9
  x = rng.randn(n, c).astype(np.float32)
  gains = np.array([1e-6, 3e-5, 1.0, 4096.0, 7e3, 1e-3] * 4, np.float32)
  x *= gains
  x[:, 6] = 0.0
  x[:, 7] = 2.5
  x[1234, 8] = 1e4
  xd = torch.as_tensor(x, device='cuda')
  stream = ops.Stream(xd, pre, post, [0, 2500], [2500, 3500])
  keep = [2500, 3400]
  design = np.concatenate([ora.add_context(x[:2500], pre, post), ora.add_context(x[2500:], pre, post)[:3400]]).astype(np.float64)
  want = design.T @ design
  lags = pre + post + 1
  for method in (4, 6, 5):
    got = ops.lagged_gram([stream], keep, options=ops.gram_options(method=method)).cpu().numpy().astype(np.float64)
    assert np.array_equal(got, got.T)
    worst = 0.0
    for c1 in range(c):
      for c2 in range(c):
        w = want[c1::c, c2::c]
        g = got[c1::c, c2::c]
        assert w.shape == (lags, lags)
        scale = np.sqrt(want[c1::c, c1::c].diagonal().mean() * want[c2::c, c2::c].diagonal().mean())
        if scale > 0:
          worst = max(worst, np.abs(g - w).max() / scale)
        else:
          assert not g.any()
    print('method %d: worst block error relative to the channels\' energy %.2e' % (method, worst))
    assert worst < GRAM_TOL, method
