"""GPU parity of the model API (fit / evaluate / predict, lambda sweep, CCA) against the fixtures
written by the unmodified reference and against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_numpy as ora
from telluride_decoding_b200 import brain_model
from telluride_decoding_b200 import cca
from telluride_decoding_b200 import decoding
from tests import gpu_util

pytestmark = pytest.mark.gpu

W_TOL = 1e-4     # BASELINE.json north_star: decoder weights and correlations


def _golden(golden_dir, name):
  return np.load(os.path.join(golden_dir, name + '.npz'))


@pytest.mark.parametrize('name', ['ridge_small', 'ridge_c8l8'])
def test_ridge_matches_reference_fixture(golden_dir, name):
  g = _golden(golden_dir, name)
  bd = gpu_util.make_test_data(g['x'], g['y'], pre_context=int(g['pre']),
                               post_context=int(g['post']), final_batch_size=int(g['batch']))
  ds = bd.create_dataset('train')
  for tag, lamb, use_offset, use_ridge in (('ridge0', 0.0, True, True), ('ridge1em2', 1e-2, True, True),
                                           ('ridge10', 10.0, True, True), ('nooffset', 0.1, False, True),
                                           ('shrink03', 0.3, True, False), ('lw', -1, True, False)):
    w, b, cov_x, cov_xy, shrink = brain_model.calculate_linear_regressor_parameters_from_dataset(
        ds, lamb=lamb, use_offset=use_offset, use_ridge=use_ridge)
    assert gpu_util.rel_fro(w, g[tag + '_w']) < W_TOL, tag
    assert gpu_util.rel_fro(cov_x, g[tag + '_cov_x']) < 1e-5, tag
    assert gpu_util.rel_fro(cov_xy, g[tag + '_cov_xy']) < 1e-5, tag
    if use_offset:
      assert np.allclose(b, g[tag + '_b'], rtol=W_TOL, atol=W_TOL * np.abs(g[tag + '_b']).max()), tag
      assert b.shape == (1, g['y'].shape[1])
    else:
      assert b.shape == (1,) and b[0] == 0
    if tag == 'lw':      # Ledoit-Wolf estimate (brain_model.py:457-465): a computed, not a passed, value
      assert abs(float(shrink) - float(g[tag + '_shrinkage'])) <= 1e-4 * abs(float(g[tag + '_shrinkage']))
    else:
      assert float(shrink) == float(g[tag + '_shrinkage'])
    assert w.dtype == np.float32
  with pytest.raises(ValueError, match='Regularization lambda must be between 0 and 1'):
    brain_model.calculate_linear_regressor_parameters_from_dataset(ds, lamb=2, use_ridge=False)
  with pytest.raises(TypeError, match='must be a tf.data.Dataset object'):
    brain_model.calculate_linear_regressor_parameters_from_dataset([1, 2, 3])


def test_linear_model_recovers_known_weights_like_reference_test():
  """test/brain_model_test.py:852-911: y = xW + b exactly -> W, b to 1e-5, eval MSE < 1e-5."""
  rng = np.random.RandomState(0)
  x = rng.randn(1000, 2).astype(np.float32)
  wt = np.array([[1, 3], [2, 4]], np.float32)
  bt = np.array([5, 6], np.float32)
  y = x @ wt + bt
  bd = gpu_util.make_test_data(x, y, final_batch_size=100)
  ds = bd.create_dataset('train')
  model = brain_model.BrainModelLinearRegression(ds, regularization_lambda=0.0)
  assert model.fit(ds) == {}
  np.testing.assert_allclose(model.w_estimate, wt, atol=1e-5)
  np.testing.assert_allclose(model.b_estimate, bt, atol=1e-5)
  w, b = model.weight_matrices
  assert w.shape == (2, 2) and b.shape == (2,)
  res = model.evaluate(bd.create_dataset('test'))
  assert set(res) == {'loss', 'pearson_correlation_first'}
  assert res['loss'] < 1e-5 and res['pearson_correlation_first'] > 0.9999
  pred = model.predict({'input_1': x[:7]})
  np.testing.assert_allclose(pred, y[:7], rtol=1e-5)
  with pytest.raises(TypeError, match='must be called with'):
    model.fit([1, 2])
  with pytest.raises(ValueError, match='Y value not needed'):
    model.fit(ds, y=3)
  with pytest.raises(TypeError, match='must be called with tf.data.Dataset'):
    model.evaluate({'input_1': x})


def test_trf_simulation_32ch_33lags():
  """test/brain_model_test.py:728-776 in spirit: 32 channels x 33 lags, r > 0.99 and weights
  equal to the oracle's."""
  rng = np.random.RandomState(4)
  n, c, pre, post = 6000, 32, 16, 16
  x = rng.randn(n, c).astype(np.float32)
  w_true = (rng.randn((pre + 1 + post) * c, 1) / 8).astype(np.float32)
  y = (ora.add_context(x, pre, post) @ w_true + 0.05 * rng.randn(n, 1)).astype(np.float32)
  bd = gpu_util.make_test_data(x, y, pre_context=pre, post_context=post, final_batch_size=500)
  ds = bd.create_dataset('train')
  model = brain_model.BrainModelLinearRegression(ds, regularization_lambda=1e-3)
  model.fit(ds)
  batches = gpu_util.oracle_batches(x, y, pre=pre, post=post, batch=500)
  w, b, _, _, _ = ora.linear_regressor_from_batches(batches, 1e-3)
  assert gpu_util.rel_fro(model.w_estimate, w) < W_TOL
  got = model.evaluate(bd.create_dataset('test'))
  want = ora.evaluate_linear(batches, w, b)
  assert got['pearson_correlation_first'] > 0.99
  assert abs(got['pearson_correlation_first'] - want['pearson_correlation_first']) < W_TOL
  assert abs(got['loss'] - want['loss']) < W_TOL * want['loss'] + 1e-7
  # predictions over the dataset come back in stream order
  pred = model.predict(bd.create_dataset('test'))
  ref = np.concatenate([np.asarray(bb[0]['input_1']) @ w + b for bb in batches], axis=0)
  np.testing.assert_allclose(pred, ref, rtol=1e-3, atol=1e-4)


def test_lambda_sweep_one_factorisation_equals_per_lambda_solves():
  rng = np.random.RandomState(8)
  n, c, post = 5000, 16, 7
  x = rng.randn(n, c).astype(np.float32)
  x[:, 1] = x[:, 0] + 0.01 * x[:, 1]        # ill-conditioned pair of channels
  y = (x[:, :1] * 0.5 + rng.randn(n, 1)).astype(np.float32)
  bd = gpu_util.make_test_data(x, y, post_context=post, final_batch_size=500)
  ds = bd.create_dataset('train')
  stats = brain_model.RegressionStats.from_dataset(ds)
  lambdas = np.logspace(-6, 3, 20)
  chol = brain_model.solve_regression(stats, lambdas, method=0)
  eigh = brain_model.solve_regression(stats, lambdas, method=1)
  sweep = brain_model.solve_regression(stats, lambdas, method=2)     # batched blocked Cholesky
  batches = gpu_util.oracle_batches(x, y, post=post, batch=500)
  for lam, (w0, b0), (w1, b1), (w2, b2) in zip(lambdas, chol, eigh, sweep):
    assert gpu_util.rel_fro(w1.cpu().numpy(), w0.cpu().numpy()) < 1e-5
    assert gpu_util.rel_fro(w2.cpu().numpy(), w0.cpu().numpy()) < 1e-6
    assert abs(float(b2[0]) - float(b0[0])) < 1e-6
    # Against the reference only where the system is well conditioned: the smallest eigenvalue of
    # this design is ~5e-5, so at lambda = 1e-3 the condition number is ~2000 and a relative
    # perturbation of 1e-7 in the Gram (the order of the fp32 red.add accumulation here, the float32
    # rounding of np.matmul and LU there) already moves the weights by up to 2e-4.
    if lam >= 0.1:
      w, b, _, _, _ = ora.linear_regressor_from_batches(batches, float(lam))
      assert gpu_util.rel_fro(w0.cpu().numpy(), w) < W_TOL, lam
      assert abs(float(b0[0]) - float(b[0, 0])) < W_TOL


def test_batched_sweep_multi_rhs_and_singular_fallback():
  """Several targets through the batched sweep; lambda = 0 on a rank-deficient design falls back to
  LU for that lambda only (np.linalg.solve semantics) or reports a singular matrix."""
  rng = np.random.RandomState(17)
  n, c = 3000, 70                 # 141 columns: panels of 64 + a ragged last panel
  x = rng.randn(n, c).astype(np.float32)
  y = (x[:, :3] @ rng.randn(3, 3) + 0.1 * rng.randn(n, 3)).astype(np.float32)
  bd = gpu_util.make_test_data(x, y, post_context=1, final_batch_size=300)
  stats = brain_model.RegressionStats.from_dataset(bd.create_dataset('train'))
  lambdas = [1e-3, 0.1, 5.0]
  a = brain_model.solve_regression(stats, lambdas, method=0)
  b = brain_model.solve_regression(stats, lambdas, method=2)
  for (w0, b0), (w2, b2) in zip(a, b):
    assert w2.shape == (140, 3)
    assert gpu_util.rel_fro(w2.cpu().numpy(), w0.cpu().numpy()) < 1e-6
    np.testing.assert_allclose(b2.cpu().numpy(), b0.cpu().numpy(), atol=1e-6)


def test_multi_trial_interleaved_fit_and_evaluate():
  """Three trials of unequal length: kept rows are a per-trial prefix of the round-robin stream
  and evaluate's batches mix trials (brain_data.py:725-728)."""
  from telluride_decoding_b200 import brain_data
  rng = np.random.RandomState(12)
  lens = [1300, 900, 1111]
  bd = brain_data.TrialBrainData('eeg', 'intensity', 100.0, pre_context=2, post_context=3,
                                 final_batch_size=200, shuffle_buffer_size=0)
  by_name = {}
  wt = rng.randn(6 * 5, 1).astype(np.float32)
  for i, n in enumerate(lens):
    x = rng.randn(n, 5).astype(np.float32)
    y = (ora.add_context(x, 2, 3) @ wt + 0.2 * rng.randn(n, 1)).astype(np.float32)
    bd.add_trial('t%d' % i, x, y)
    z = np.zeros((n, 1), np.float32)
    by_name['t%d' % i] = ora.file_rows(x, z, y, z, 2, 3)
  ds = bd.create_dataset('train')
  # like the reference (brain_data.py:206-207) the file list is shuffled once; the interleave
  # follows that order
  files = [by_name[name] for name, _, _ in ds.files]
  assert ds.num_batches == sum(lens) // 200
  model = brain_model.BrainModelLinearRegression(ds, regularization_lambda=0.1)
  model.fit(ds)
  batches = ora.batches_from_files(files, 200)
  w, b, _, _, _ = ora.linear_regressor_from_batches(batches, 0.1)
  assert gpu_util.rel_fro(model.w_estimate, w) < W_TOL
  got = model.evaluate(bd.create_dataset('test'))
  want = ora.evaluate_linear(batches, w, b)
  assert abs(got['pearson_correlation_first'] - want['pearson_correlation_first']) < W_TOL
  # iteration reproduces the interleaved minibatches bit for bit
  for (gd, gy), (od, oy) in zip(bd.create_dataset('program_test'), batches):
    np.testing.assert_array_equal(gd['input_1'].cpu().numpy(), od['input_1'])
    np.testing.assert_array_equal(gy.cpu().numpy(), oy)


def _align(rot, ref):
  """Joint sign per component is arbitrary in the reference (SVD): align before comparing."""
  s = np.sign(np.sum(rot * ref, axis=0))
  assert np.all(s != 0)
  return rot * s


def _cca_compare(rot_x, rot_y, e, g, twin):
  """North-star contract: 1e-4 on the rotations and correlations.  The reference computes in the
  float32 of its inputs (general `eig` twice + SVD, cca.py:345-362), which leaves rounding of that
  order in ITS output, so the comparison has two parts: (i) against the float64 twin of the
  reference's algorithm we are held to 1e-5; (ii) against the reference's own float32
  output to 1e-4 plus the reference's measured distance from its twin."""
  tx, ty, _, _, te = twin
  for i in range(rot_x.shape[1]):
    gx, gy = g['rot_x'][:, i:i + 1], g['rot_y'][:, i:i + 1]
    ax, ay = _align(rot_x[:, i:i + 1], gx), _align(rot_y[:, i:i + 1], gy)
    wx, wy = _align(tx[:, i:i + 1], gx), _align(ty[:, i:i + 1], gy)
    assert gpu_util.rel_fro(ax, wx) < 1e-5, i
    assert gpu_util.rel_fro(ay, wy) < 1e-5, i
    ref_noise = max(gpu_util.rel_fro(gx, wx), gpu_util.rel_fro(gy, wy))
    assert gpu_util.rel_fro(ax, gx) < W_TOL + ref_noise, (i, ref_noise)
    assert gpu_util.rel_fro(ay, gy) < W_TOL + ref_noise, (i, ref_noise)
  np.testing.assert_allclose(e, te, rtol=1e-5)
  np.testing.assert_allclose(e, g['e'], rtol=W_TOL)


def test_cca_matches_reference_fixture(golden_dir):
  g = _golden(golden_dir, 'cca_small')
  n = g['x'].shape[0]
  bd = gpu_util.make_test_data(g['x'], np.ones((n, 1), np.float32), g['x2'],
                               post_context=int(g['post1']), in2_post_context=int(g['post2']),
                               final_batch_size=int(g['batch']))
  ds = bd.create_dataset('train')
  rot_x, rot_y, mean_x, mean_y, e = cca.calculate_cca_parameters_from_dataset(
      ds, int(g['dim']), regularization=float(g['reg']), mini_batch_count=0)
  batches = gpu_util.oracle_batches(g['x'], np.ones((n, 1), np.float32), g['x2'], None, 0,
                                    int(g['post1']), 0, int(g['post2']), 0, int(g['batch']))
  twin = ora.cca_from_batches(batches, int(g['dim']), float(g['reg']), mini_batch_count=0,
                              dtype=np.float64)
  _cca_compare(rot_x, rot_y, e, g, twin)
  np.testing.assert_allclose(mean_x, g['mean_x'], atol=1e-6)
  np.testing.assert_allclose(mean_y, g['mean_y'], atol=1e-6)
  assert rot_x.dtype == np.float32 and mean_x.shape == (1, rot_x.shape[0])
  # model API + evaluate against the oracle forward / per-batch Pearson
  model = cca.BrainModelCCA(ds, cca_dims=int(g['dim']), regularization_lambda=float(g['reg']))
  assert model.fit(ds) == {}
  res = model.evaluate(bd.create_dataset('test'))
  want = ora.evaluate_cca(batches, g['mean_x'], g['mean_y'], g['rot_x'], g['rot_y'])
  assert set(res) == {'loss', 'cca_pearson_correlation_first'}
  assert abs(res['cca_pearson_correlation_first'] - want['cca_pearson_correlation_first']) < W_TOL
  assert res['loss'] == res['cca_pearson_correlation_first']
  out = model.predict({'input_1': batches[0][0]['input_1'], 'input_2': batches[0][0]['input_2']})
  assert out.shape == (int(g['batch']), 2 * int(g['dim']))
  with pytest.raises(ValueError, match='regularization lambda must be >= 0'):
    cca.calculate_cca_parameters_from_dataset(ds, 3, regularization=-1)
  with pytest.raises(TypeError, match='must be a tf.data.Dataset object'):
    cca.calculate_cca_parameters_from_dataset([1], 3)


def test_cca_cholesky_whitening_equals_the_eigendecomposition_route(golden_dir):
  """td_cca_solve takes Cholesky whitening when reg > 10 * eps_eig and the reference's route
  (syevd, eps threshold, gesvd) otherwise: at reg = 0.1 an eps_eig of 0.02 drops no eigenvalue
  (all are >= reg) but selects the second route, so both must give the same canonical pairs."""
  g = _golden(golden_dir, 'cca_small')
  bd = gpu_util.make_test_data(g['x'], np.ones((g['x'].shape[0], 1), np.float32), g['x2'],
                               post_context=int(g['post1']), in2_post_context=int(g['post2']),
                               final_batch_size=int(g['batch']))
  stats = cca.CcaStats.from_dataset(bd.create_dataset('train'))
  dim = int(g['dim'])
  fast = cca.solve_cca(stats, dim, 0.1, eps_eig=1e-12)
  slow = cca.solve_cca(stats, dim, 0.1, eps_eig=0.02)
  for a, b in zip(fast, slow):
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-5, atol=2e-6)
  np.testing.assert_allclose(fast[4].cpu().numpy().reshape(-1), g['e'].reshape(-1)[:dim], rtol=1e-4)


def test_cca_seed42_structure_like_reference_test(golden_dir):
  """test/cca_test.py:80-123 construction: eigenvalue pattern and the reference's own output."""
  g = _golden(golden_dir, 'cca_seed42')
  bd = gpu_util.make_test_data(g['x'], np.ones((g['x'].shape[0], 1), np.float32), g['y'],
                               final_batch_size=100)
  ds = bd.create_dataset('train')
  rot_x, rot_y, _, _, e = cca.calculate_cca_parameters_from_dataset(ds, 3, regularization=0.1,
                                                                     mini_batch_count=0)
  assert e[0] > 0.9 and e[1] > 0.6 and e[2] < 0.1
  np.testing.assert_allclose(e, g['e'], rtol=W_TOL)
  # mini_batch_count limits the statistics to the first batches of the stream (cca.py:311)
  first5 = cca.calculate_cca_parameters_from_dataset(ds, 3, 0.1, mini_batch_count=5)
  batches = [({'input_1': g['x'][i:i + 100], 'input_2': g['y'][i:i + 100]}, None)
             for i in range(0, 2000, 100)]
  want = ora.cca_from_batches(batches, 3, 0.1, mini_batch_count=5)
  np.testing.assert_allclose(first5[4], want[4], rtol=W_TOL)


def test_cca_width_errors_like_reference_test():
  """test/cca_test.py:199-220."""
  x = np.random.randn(200, 1).astype(np.float32)
  bd = gpu_util.make_test_data(x, x, x, final_batch_size=50)
  with pytest.raises(ValueError, match=r'Input 1 feature width \(1\) should not be <= 1'):
    cca.BrainModelCCA(bd.create_dataset('train'))


def test_create_brain_model_and_train_and_test():
  rng = np.random.RandomState(2)
  x = rng.randn(2000, 4).astype(np.float32)
  y = (x[:, :1] + 0.1 * rng.randn(2000, 1)).astype(np.float32)
  bd = gpu_util.make_test_data(x, y, post_context=2, final_batch_size=100)
  opts = decoding.DecodingOptions(dnn_regressor='linear', regularization_lambda=0.01)
  model = decoding.create_brain_model(opts, bd.create_dataset('train'))
  train_res, test_res = decoding.train_and_test(opts, bd, model)
  assert train_res == {} and test_res['pearson_correlation_first'] > 0.99
  with pytest.raises(TypeError, match='outside the B200 dense decoding path'):
    decoding.create_brain_model(decoding.DecodingOptions(), bd.create_dataset('train'))


def test_pearson_known_answer():
  """test/brain_model_test.py:1046-1082."""
  d = np.array([[1, 43, 99], [2, 21, 65], [3, 25, 79], [4, 42, 75], [5, 57, 87], [6, 59, 81]],
               dtype=np.float32)
  r = brain_model.pearson_correlation(torch.as_tensor(d[:, 1:2], device='cuda'),
                                      torch.as_tensor(d[:, 2:3], device='cuda'))
  assert abs(float(r[0]) - 0.5298) < 1e-4
  assert abs(float(brain_model.pearson_correlation_first(d[:, 1:2], d[:, 2:3])) - 0.5298) < 1e-4
  x = np.random.random((6, 2))
  y = np.random.random((6, 2))
  npcor = np.diag(np.corrcoef(x, y, rowvar=False)[:2, 2:])
  np.testing.assert_allclose(brain_model.pearson_correlation(x, y).numpy(), npcor, atol=2e-7)


def test_pinned_host_trials_are_uploaded_in_chunks_and_give_the_same_fit():
  """TrialBrainData.set_trials with pinned host arrays: asynchronous chunked upload, statistics per
  arrived group of trials -- same numbers as with resident tensors, and every other consumer of the
  dataset (iteration, evaluate) waits for the copies."""
  import torch
  from telluride_decoding_b200 import brain_data
  rng = np.random.default_rng(21)
  lengths = [700, 650, 900, 512, 801, 640, 777]
  n = sum(lengths)
  x = rng.standard_normal((n, 12)).astype(np.float32)
  y = (x[:, :1] * 0.7 + 0.2 * rng.standard_normal((n, 1))).astype(np.float32)
  names = ['trial_%d' % i for i in range(len(lengths))]

  def make(xa, ya):
    bd = brain_data.TrialBrainData('eeg', 'envelope', 64.0, pre_context=1, post_context=6,
                                   final_batch_size=100, shuffle_buffer_size=0)
    bd.set_trials(names, lengths, xa, ya)
    return bd

  random_state = __import__('random').getstate()
  resident = make(torch.as_tensor(x, device='cuda'), torch.as_tensor(y, device='cuda'))
  __import__('random').setstate(random_state)           # same shuffled file order
  hosted = make(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory())
  assert hosted._ready is not None and len(hosted._ready) > 1
  ds_r, ds_h = resident.create_dataset('train'), hosted.create_dataset('train')
  assert len(ds_h.segment_groups()) > 1
  s_r = brain_model.RegressionStats.from_dataset(ds_r)
  s_h = brain_model.RegressionStats.from_dataset(ds_h)
  assert gpu_util.rel_fro(s_h.gram.cpu().numpy(), s_r.gram.cpu().numpy()) < 1e-6
  np.testing.assert_allclose(s_h.xtv.cpu().numpy(), s_r.xtv.cpu().numpy(), rtol=1e-9, atol=1e-9)
  assert s_h.n == s_r.n
  # a fresh hosted object consumed by iteration first
  __import__('random').setstate(random_state)
  hosted2 = make(torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory())
  for (da, ya), (db, yb) in zip(hosted2.create_dataset('train'), ds_r):
    np.testing.assert_array_equal(da['input_1'].cpu().numpy(), db['input_1'].cpu().numpy())
    np.testing.assert_array_equal(ya.cpu().numpy(), yb.cpu().numpy())
  model = brain_model.BrainModelLinearRegression(ds_h, regularization_lambda=0.1)
  model.fit(hosted2.create_dataset('train'))
  ref = brain_model.BrainModelLinearRegression(ds_r, regularization_lambda=0.1)
  ref.fit(ds_r)
  assert gpu_util.rel_fro(model.weight_matrices[0], ref.weight_matrices[0]) < 1e-5


def test_inputs_cross_by_dlpack_and_shutdown_releases_handles():
  """north_star: tensors cross the boundary by DLPack.  A non-torch producer (an object exposing only
  `__dlpack__` / `__dlpack_device__`) and a raw DLPack capsule reach the kernels without a host
  copy and give the same fit as the numpy path; td_shutdown releases the library's handles and the
  next call recreates them."""
  from telluride_decoding_b200 import _lib, brain_data

  class Producer(object):          # what cupy / jax arrays look like to a consumer
    def __init__(self, t):
      self._t = t

    def __dlpack__(self, stream=None):
      return self._t.__dlpack__(stream=stream)

    def __dlpack_device__(self):
      return self._t.__dlpack_device__()

  rng = np.random.RandomState(3)
  x = rng.randn(2000, 6).astype(np.float32)
  y = (x[:, :1] - 0.5 * x[:, 2:3] + 0.1 * rng.randn(2000, 1)).astype(np.float32)
  xd, yd = torch.as_tensor(x, device='cuda'), torch.as_tensor(y, device='cuda')
  got = _lib.as_device_f32(Producer(xd))
  assert got.data_ptr() == xd.data_ptr()              # zero copy
  cap = torch.utils.dlpack.to_dlpack(yd)
  assert _lib.as_device_f32(cap).data_ptr() == yd.data_ptr()
  weights = []
  for xin, yin in ((x, y), (Producer(xd), torch.utils.dlpack.to_dlpack(yd))):
    bd = brain_data.TestBrainData('input_1', 'output', 100.0, pre_context=1, post_context=2,
                                  final_batch_size=100, shuffle_buffer_size=0)
    bd.preserve_test_data(xin, yin)
    ds = bd.create_dataset('train')
    model = brain_model.BrainModelLinearRegression(ds, regularization_lambda=0.01)
    model.fit(ds)
    weights.append(model.w_estimate.copy())
    assert _lib.load().td_shutdown() == 0             # handles are recreated on the next solve
  # (the Gram's partial tiles are combined with fp32 atomics, so two runs differ in the last bits)
  np.testing.assert_allclose(weights[0], weights[1], rtol=1e-4, atol=1e-6)


def test_long_streams_with_small_batches_and_unit_windows():
  """The windowed-correlation kernel keeps its per-chunk partial sums in global memory: no limit on
  segment length x columns (the round-1 kernel held them in 192 KB of shared memory and refused
  evaluate() at the regression preset's batch_size = 100 beyond ~98 k rows with 5 CCA dimensions, and
  Decoder.train(window_size=0) beyond ~4.9 k rows).  cfg2 length: 460 800 rows."""
  from telluride_decoding_b200 import infer_decoder
  rng = np.random.RandomState(8)
  n = 460800
  x = rng.randn(n, 4).astype(np.float32)
  y = (x[:, :1] + 0.5 * rng.randn(n, 1)).astype(np.float32)
  bd = gpu_util.make_test_data(x, y, post_context=3, final_batch_size=100)
  ds = bd.create_dataset('train')
  model = brain_model.BrainModelLinearRegression(ds, regularization_lambda=0.1)
  model.fit(ds)
  res = model.evaluate(bd.create_dataset('test'))
  # Keras semantics: mean over the 4608 minibatches of the per-batch Pearson r
  pred = model.predict(bd.create_dataset('test'))
  r = [np.corrcoef(pred[i:i + 100, 0], y[i:i + 100, 0])[0, 1] for i in range(0, n, 100)]
  assert abs(res['pearson_correlation_first'] - np.mean(r)) < 1e-5
  # unit windows over five columns: the per-frame signal itself, 2.3 M values
  a = torch.as_tensor(rng.randn(n, 5).astype(np.float32), device='cuda')
  b = torch.as_tensor(rng.randn(n, 5).astype(np.float32), device='cuda')
  mu = torch.zeros(5, device='cuda'); pw = torch.full((5,), 2.0, device='cuda')
  from telluride_decoding_b200 import ops
  vals, _ = ops.windowed_corr(a, b, [0], [n], [1], [1], 'global_all', mu_a=mu, mu_b=mu, power=pw)
  want = (a * b / 2.0).double()
  assert torch.equal(vals, want)


def test_ledoit_wolf_shrinkage_on_interleaved_trials():
  """lamb = -1, use_ridge=False (brain_model.py:440-443, :457-465): the running-mean statistic comes
  from two kernels over the implicit design; minibatches of the interleaved stream mix the trials."""
  from telluride_decoding_b200 import brain_data
  rng = np.random.RandomState(14)
  lens = [1300, 900, 1111]
  bd = brain_data.TrialBrainData('eeg', 'intensity', 100.0, pre_context=2, post_context=3,
                                 final_batch_size=200, shuffle_buffer_size=0)
  by_name = {}
  wt = rng.randn(6 * 5, 1).astype(np.float32)
  for i, n in enumerate(lens):
    x = (rng.randn(n, 5) * (1 + i) + 0.3).astype(np.float32)
    y = (ora.add_context(x, 2, 3) @ wt + 0.2 * rng.randn(n, 1)).astype(np.float32)
    bd.add_trial('t%d' % i, x, y)
    z = np.zeros((n, 1), np.float32)
    by_name['t%d' % i] = ora.file_rows(x, z, y, z, 2, 3)
  ds = bd.create_dataset('test')
  batches = ora.batches_from_files([by_name[name] for name, _, _ in ds.files], 200)
  for use_offset in (True, False):
    w, b, _, _, shrink = brain_model.calculate_linear_regressor_parameters_from_dataset(
        ds, lamb=-1, use_offset=use_offset, use_ridge=False)
    w0, b0, _, _, shrink0 = ora.linear_regressor_from_batches(batches, -1, use_offset=use_offset,
                                                              use_ridge=False)
    assert abs(shrink - shrink0) < 1e-4 * max(abs(shrink0), 1e-3), (shrink, shrink0)
    assert gpu_util.rel_fro(w, w0) < W_TOL
  # pieces of the minibatches: every row of the kept stream exactly once, 200 rows per minibatch
  pieces = ds.batch_pieces()
  per_batch = {}
  for s, lo, hi, batch in pieces:
    per_batch[batch] = per_batch.get(batch, 0) + hi - lo
  assert set(per_batch.values()) == {200} and len(per_batch) == ds.num_batches


def test_lambda_sweep_matches_the_reference_at_every_lambda_on_a_bench_like_design():
  """The bench's sweep (logspace(-6, 3, 20)) on a design like the bench's (z-scored white sensors, an
  FIR-driven envelope): weights within 1e-4 of the reference at ALL 20 lambdas -- here the design is
  well conditioned, so the reference's own float32 solve is trustworthy down to lambda = 1e-6."""
  rng = np.random.RandomState(21)
  n, c, post = 23040, 16, 15                      # 3 trials' worth of frames, D = 256
  x = rng.randn(n, c).astype(np.float32)
  taps = (rng.randn(10, c) / 10).astype(np.float32)
  y = np.zeros((n, 1), np.float32)
  for k in range(10):
    y[:n - k, 0] += x[k:] @ taps[k]
  y += 0.5 * rng.randn(n, 1).astype(np.float32)
  bd = gpu_util.make_test_data(x, y, post_context=post, final_batch_size=768)
  stats = brain_model.RegressionStats.from_dataset(bd.create_dataset('train'))
  lambdas = np.logspace(-6, 3, 20)
  sweep = brain_model.solve_regression(stats, lambdas, method=2)
  batches = gpu_util.oracle_batches(x, y, post=post, batch=768)
  for lam, (w2, b2) in zip(lambdas, sweep):
    w, b, _, _, _ = ora.linear_regressor_from_batches(batches, float(lam))
    assert gpu_util.rel_fro(w2.cpu().numpy(), w) < W_TOL, lam
    assert abs(float(b2[0]) - float(b[0, 0])) < W_TOL, lam
