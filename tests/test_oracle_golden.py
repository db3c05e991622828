"""Pins oracle/ref_numpy.py against fixtures written by the UNMODIFIED reference
(oracle/gen_golden.py) and against the literal known-answer vectors in the reference's
own unit tests.  CPU only."""
import os

import numpy as np
import pytest

from oracle import ref_numpy as ora


def _load(golden_dir, name):
  return np.load(os.path.join(golden_dir, name + '.npz'))


def _batches(g):
  x, y = g['x'], g['y']
  zeros = np.zeros((x.shape[0], 1), np.float32)
  rows = ora.file_rows(x, zeros, y, zeros, int(g['pre']), int(g['post']))
  return ora.batches_from_files([rows], int(g['batch']))


@pytest.mark.parametrize('name', ['ridge_small', 'ridge_c8l8'])
@pytest.mark.parametrize('tag,lamb,use_offset,use_ridge', [
    ('ridge0', 0.0, True, True), ('ridge1em2', 1e-2, True, True),
    ('ridge10', 10.0, True, True), ('nooffset', 0.1, False, True),
    ('shrink03', 0.3, True, False), ('lw', -1, True, False)])
def test_ridge_matches_reference(golden_dir, name, tag, lamb, use_offset, use_ridge):
  g = _load(golden_dir, name)
  w, b, cov_x, cov_xy, shrink = ora.linear_regressor_from_batches(
      _batches(g), lamb, use_offset, use_ridge)
  # Same arithmetic, same BLAS: bit-identical.
  np.testing.assert_array_equal(w, g[tag + '_w'])
  np.testing.assert_array_equal(np.asarray(b), g[tag + '_b'])
  np.testing.assert_array_equal(cov_x, g[tag + '_cov_x'])
  np.testing.assert_array_equal(cov_xy, g[tag + '_cov_xy'])
  assert float(shrink) == float(g[tag + '_shrinkage'])
  if use_ridge:
    assert w.dtype == np.float32    # in-place += of a float64 identity keeps float32


def test_cca_matches_reference(golden_dir):
  g = _load(golden_dir, 'cca_small')
  n = g['x'].shape[0]
  rows = ora.file_rows(g['x'], g['x2'], np.ones((n, 1), np.float32),
                       np.zeros((n, 1), np.float32), 0, int(g['post1']), 0, int(g['post2']))
  batches = ora.batches_from_files([rows], int(g['batch']))
  rot_x, rot_y, mean_x, mean_y, e = ora.cca_from_batches(
      batches, int(g['dim']), float(g['reg']), mini_batch_count=0)
  np.testing.assert_array_equal(rot_x, g['rot_x'])
  np.testing.assert_array_equal(rot_y, g['rot_y'])
  np.testing.assert_array_equal(mean_x, g['mean_x'])
  np.testing.assert_array_equal(e, g['e'])


def test_cca_seed42_structure(golden_dir):
  """test/cca_test.py:121-123 style thresholds on the reference's own output, and the
  restatement equal to it."""
  g = _load(golden_dir, 'cca_seed42')
  x, y = g['x'], g['y']
  batches = [({'input_1': x[i:i + 100], 'input_2': y[i:i + 100]}, None)
             for i in range(0, x.shape[0], 100)]
  rot_x, rot_y, _, _, e = ora.cca_from_batches(batches, 3, 0.1, mini_batch_count=0)
  np.testing.assert_array_equal(e, g['e'])
  np.testing.assert_array_equal(rot_x, g['rot_x'])
  assert e[0] > 0.9 and e[1] > 0.6 and e[2] < 0.1


def test_lag_goldens_from_reference_tests():
  """test/brain_data_test.py:291-294 (pre=2) and :317-320 (post=2): two channels
  [t, 1000+t]."""
  t = np.arange(20, dtype=np.float32)
  x = np.stack([t, 1000 + t], axis=1)
  pre2 = ora.add_context(x, 2, 0)
  np.testing.assert_array_equal(pre2[:3], [[0, 0, 0, 0, 0, 1000],
                                           [0, 0, 0, 1000, 1, 1001],
                                           [0, 1000, 1, 1001, 2, 1002]])
  post2 = ora.add_context(x, 0, 2)
  np.testing.assert_array_equal(post2[:3], [[0, 1000, 1, 1001, 2, 1002],
                                            [1, 1001, 2, 1002, 3, 1003],
                                            [2, 1002, 3, 1003, 4, 1004]])
  np.testing.assert_array_equal(post2[-1], [19, 1019, 0, 0, 0, 0])
  assert pre2.shape == (20, 6)


def test_input_offset_goldens():
  """test/brain_data_test.py:232-272."""
  t = np.arange(30, dtype=np.float32)
  x = np.stack([t, 1000 + t], axis=1)
  x2 = (2000 + t).reshape(-1, 1)
  y = (3000 + t).reshape(-1, 1)
  a = np.zeros((30, 1), np.float32)
  for off, x0, x20, y0 in ((1, [1, 1001], 2000, 3000), (-1, [0, 1000], 2001, 3001),
                           (2, [2, 1002], 2000, 3000)):
    xr, x2r, yr, _ = ora.file_rows(x, x2, y, a, input_offset=off)
    np.testing.assert_array_equal(xr[0], x0)
    assert x2r[0, 0] == x20 and yr[0, 0] == y0
    assert xr.shape[0] == 30 - abs(off)


def test_chunk_edges_are_zero_padded():
  x = np.arange(1, 11, dtype=np.float32).reshape(-1, 1)
  z = np.zeros((10, 1), np.float32)
  rows = ora.file_rows(x, z, z, z, pre=1, post=1, initial_batch_size=5)[0]
  np.testing.assert_array_equal(rows[4], [4, 5, 0])    # end of chunk 0
  np.testing.assert_array_equal(rows[5], [0, 6, 7])    # start of chunk 1


def test_interleave_and_drop_remainder():
  fi, ri = ora.interleave_order([3, 1, 2])
  np.testing.assert_array_equal(fi, [0, 1, 2, 0, 2, 0])
  np.testing.assert_array_equal(ri, [0, 0, 0, 1, 1, 2])
  files = []
  for f, n in enumerate([3, 1, 2]):
    v = (100 * f + np.arange(n, dtype=np.float32)).reshape(-1, 1)
    files.append((v, v, v, v))
  b = ora.batches_from_files(files, 4)
  assert len(b) == 1
  np.testing.assert_array_equal(b[0][1][:, 0], [0, 100, 200, 1])


def test_pearson_known_answer():
  """test/brain_model_test.py:1046-1065: r = 0.5298 +- 1e-4."""
  d = np.array([[1, 43, 99], [2, 21, 65], [3, 25, 79], [4, 42, 75], [5, 57, 87],
                [6, 59, 81]], dtype=np.float32)
  r = ora.pearson_correlation(d[:, 1:2], d[:, 2:3])
  assert abs(float(r[0]) - 0.5298) < 1e-4
  rng = np.random.RandomState(0)
  x, y = rng.random_sample((6, 2)), rng.random_sample((6, 2))
  ref = np.diag(np.corrcoef(x, y, rowvar=False)[:2, 2:])
  np.testing.assert_allclose(ora.pearson_correlation(x, y), ref, atol=2e-7)
  assert np.all(ora.pearson_correlation(np.ones((5, 2)), y[:5]) == 0)


def test_decoder_pipeline_matches_reference(golden_dir):
  g = _load(golden_dir, 'decoder')
  truth, label, batch = g['truth'], g['label'], int(g['batch'])
  results = {}
  for which, p in (('s1', g['pred']), ('s2', g['pred2'])):
    cor = ora.Correlator()
    for s in range(0, truth.shape[0], batch):
      cor.add(truth[s:s + batch], p[s:s + batch])
    np.testing.assert_array_equal(cor.mean_x, g[which + '_mean_x'])
    np.testing.assert_array_equal(cor.power, g[which + '_power'])
    c = cor.correlate(truth, p)
    np.testing.assert_array_equal(c, g[which + '_corr'])
    assert c.dtype == np.float32
    for red in ('first', 'second', 'mean', 'mean-squared'):
      sig = ora.reduce_correlations(c, red)
      for width in (10, 100, 200, 400, 700, 1000):
        res, lab = ora.windowed_means(sig, label, width)
        np.testing.assert_allclose(res, g['%s_%s_w%d_res' % (which, red, width)],
                                   rtol=0, atol=1e-15)
        np.testing.assert_array_equal(lab, g['%s_%s_w%d_lab' % (which, red, width)])
        results[(which, red, width)] = res
  for width in (10, 100, 200, 400, 700, 1000):
    r1, r2 = results[('s1', 'first', width)], results[('s2', 'first', width)]
    np.testing.assert_array_equal(ora.wta_attention(r1, r2), g['wta_w%d' % width])
    np.testing.assert_array_equal(ora.step_attention(r1, r2), g['step_w%d' % width])
  np.testing.assert_allclose(ora.average_data(g['s1_corr'], 7), g['average_data_w7'],
                             rtol=1e-6)
  assert abs(ora.calculate_dprime(g['s2_corr'][:, 0], g['s1_corr'][:, 0]) -
             float(g['dprime'])) < 1e-12


def test_window_indexing_goldens(golden_dir):
  """test/result_store_test.py:198-218: (N-W)//S+1 windows; contents start at i*S."""
  g = _load(golden_dir, 'decoder')
  starts = ora.window_starts(1000, 201, 100)
  assert len(starts) == 8 == g['store_201_100'].shape[0]
  np.testing.assert_array_equal(starts, g['store_201_100'][:, 0])
  np.testing.assert_array_equal(starts + 200, g['store_201_100'][:, 1])
  assert len(ora.window_starts(100, 201, 100)) == 0


def test_wta_and_step_goldens_from_reference_tests():
  """test/attention_decoder_test.py:111-150 literal sequences (the step decoder's
  0.5000000000000001 > 0.5 floating-point quirk included)."""
  cor2 = np.ones(13)
  cor1 = np.array([2] * 7 + [0] * 6, dtype=float)
  np.testing.assert_array_equal(ora.wta_attention(cor1, cor2), [1] * 7 + [0] * 6)
  np.testing.assert_array_equal(ora.step_attention(cor1, cor2), [1] * 11 + [0] * 2)
  cor1 = np.array([2] * 2 + [0] * 11, dtype=float)
  np.testing.assert_array_equal(ora.step_attention(cor1, cor2), [1] * 3 + [0] * 10)
  assert ora.wta_attention(0.6, 0.4) and not ora.wta_attention(0.4, 0.6)
  assert ora.aad_accuracy([True, False, True], [0, 1, 0]) == 1.0


def test_streaming_context_matches_reference_preprocessor(golden_dir):
  """Preprocessor.add_context fed batch by batch (preprocess.py:487-522; the reference's own check
  is test/preprocess_test.py:235-251)."""
  g = np.load(os.path.join(golden_dir, 'streaming_context.npz'))
  x = g['x']
  for tag in ('p2q3', 'p0q4', 'p3q0'):
    pre, post = [int(v) for v in g[tag + '_ctx']]
    cuts = [int(v) for v in g[tag + '_cuts']]
    ctx = ora.StreamingContext(pre, post)
    outs = []
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
      got = ctx.add_context(x[a:b])
      np.testing.assert_array_equal(got, g['%s_%d' % (tag, i)].astype(np.float32))
      outs.append(got)
    # the stream of rows equals the offline lag builder, delayed by post_context frames
    full = ora.add_context(x, pre, post)
    np.testing.assert_array_equal(np.concatenate(outs), full[:x.shape[0] - post])
