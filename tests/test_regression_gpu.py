"""GPU parity of the single-pass lambda-sweep x leave-one-trial-out jackknife against the
reference's loop structure (one full fit per (lambda, fold)) executed with the oracle."""
import numpy as np
import pytest

from oracle import ref_numpy as ora
from telluride_decoding_b200 import brain_data
from telluride_decoding_b200 import decoding
from telluride_decoding_b200 import regression

pytestmark = pytest.mark.gpu


def _trials(seed, lens, c, lags, c2=0):
  rng = np.random.RandomState(seed)
  wt = (rng.randn(lags * c, 1) / 3).astype(np.float32)
  out = []
  for n in lens:
    x = rng.randn(n, c).astype(np.float32)
    y = (ora.add_context(x, 0, lags - 1) @ wt + 0.5 * rng.randn(n, 1)).astype(np.float32)
    x2 = None
    if c2:
      x2 = (np.repeat(y, c2, axis=1) + rng.randn(n, c2)).astype(np.float32)
    out.append((x, y, x2))
  return out


def _oracle_jackknife(trials, names, lambdas, post, batch, cca_dims=0, post2=0, file_order=None,
                      cca_dtype=None):
  res = np.zeros((len(lambdas), len(names)))
  file_order = [names.index(n) for n in (file_order or names)]   # interleave order of the data object
  rows = []
  for x, y, x2 in trials:
    z = np.zeros((x.shape[0], 1), np.float32)
    rows.append(ora.file_rows(x, z if x2 is None else x2, y, z, 0, post, 0, post2))
  order = np.argsort(names)
  for j, f in enumerate(order):
    train = ora.batches_from_files([rows[i] for i in file_order if i != f], batch)
    test = ora.batches_from_files([rows[f]], batch)
    for i, lam in enumerate(lambdas):
      if cca_dims:
        rx, ry, mx, my, _ = ora.cca_from_batches(train, cca_dims, lam, mini_batch_count=0, dtype=cca_dtype)
        res[i, j] = ora.evaluate_cca(test, mx, my, rx, ry)['cca_pearson_correlation_first']
      else:
        w, b, _, _, _ = ora.linear_regressor_from_batches(train, lam)
        res[i, j] = ora.evaluate_linear(test, w, b)['pearson_correlation_first']
  return res


def test_linear_jackknife_matches_reference_loop():
  lens = [1230, 1111, 1404, 987, 1300]          # totals not divisible by the batch size
  names = ['trial_%02d' % i for i in range(len(lens))]
  trials = _trials(7, lens, 6, 5)
  opts = decoding.DecodingOptions()
  reg = regression.RegressionLinear(opts)
  reg.preset_flags()
  opts.post_context = 4
  bd = brain_data.TrialBrainData('eeg', 'intensity', 100.0, post_context=4,
                                 final_batch_size=opts.batch_size,
                                 shuffle_buffer_size=opts.shuffle_buffer_size)
  for n, (x, y, _) in zip(names, trials):
    bd.add_trial(n, x, y)
  lambdas = [1e-4, 1e-2, 1.0, 10.0]
  out = reg.jackknife_over_regularizations(opts, lambdas, brain_data_object=bd)
  want = _oracle_jackknife(trials, names, lambdas, 4, 100, file_order=bd.all_files())
  got = reg.all_runs_results
  assert got.shape == want.shape == (4, 5)
  np.testing.assert_allclose(got, want, atol=1e-4)
  for lam, row in zip(lambdas, want):
    mean, std = out[lam]
    assert abs(mean - row.mean()) < 1e-4 and abs(std - row.std()) < 1e-4
  one = regression.jackknife_one_model(bd, None, None, opts.set_from_dict(
      {'regularization_lambda': 1e-2}))
  np.testing.assert_allclose(one, want[1], atol=1e-4)


def test_cca_jackknife_matches_reference_loop():
  lens = [1500, 1320, 1417]
  names = ['s%d' % i for i in range(3)]
  trials = _trials(9, lens, 5, 4, c2=3)
  opts = decoding.DecodingOptions(dnn_regressor='cca', cca_dimensions=2, post_context=3,
                                  input2_post_context=2, batch_size=100, shuffle_buffer_size=0,
                                  test_metric='cca_pearson_correlation_first',
                                  train_file_pattern='allbut')
  bd = brain_data.TrialBrainData('eeg', 'eeg', 100.0, post_context=3, in2_fields='intensity',
                                 in2_post_context=2, final_batch_size=100, shuffle_buffer_size=0)
  for n, (x, y, x2) in zip(names, trials):
    bd.add_trial(n, x, y, x2)
  lambdas = [0.01, 0.1, 1.0]
  got = regression.jackknife_sweep(bd, opts, lambdas)
  want = _oracle_jackknife(trials, names, lambdas, 3, 100, cca_dims=2, post2=2,
                           file_order=bd.all_files())
  # the float64 twin of the reference's algorithm says how much of a difference is the
  # reference's own float32 eig / SVD rounding: we are held to 1e-4 against the twin, and to 1e-4 plus
  # that rounding against the float32 loop
  twin = _oracle_jackknife(trials, names, lambdas, 3, 100, cca_dims=2, post2=2,
                           file_order=bd.all_files(), cca_dtype=np.float64)
  np.testing.assert_allclose(got, twin, atol=1e-4)
  assert np.all(np.abs(got - want) <= 1e-4 + np.abs(want - twin))
