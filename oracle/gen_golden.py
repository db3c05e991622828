"""Generate tests/golden/*.npz from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference is mounted):

    python oracle/gen_golden.py

The reference's numpy hot-path functions are imported as they are from
/root/reference/telluride_decoding (TensorFlow replaced by oracle/tfstub, which only
supplies a list-backed Dataset) and executed on seeded inputs; inputs and outputs are
stored as small .npz fixtures.  The GPU box never sees /root/reference: tests there
read only the committed fixtures.
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'tfstub'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np                                            # noqa: E402
import tensorflow as tf                                       # noqa: E402  (the stub)
from telluride_decoding import attention_decoder              # noqa: E402
from telluride_decoding import brain_model                    # noqa: E402
from telluride_decoding import cca                            # noqa: E402
from telluride_decoding import infer_decoder                  # noqa: E402
from telluride_decoding import preprocess                     # noqa: E402
from telluride_decoding import result_store                   # noqa: E402
from telluride_decoding import scaled_lda                     # noqa: E402

from oracle import ref_numpy as ora                           # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _dataset(batches):
  return tf.data.Dataset(batches)


def ridge_case(name, seed, n_frames, channels, pre, post, dy, batch, noise):
  """Reference ridge / shrinkage on a lagged design built by the restated lag builder."""
  rng = np.random.RandomState(seed)
  x = rng.randn(n_frames, channels).astype(np.float32)
  lags = pre + 1 + post
  w_true = rng.randn(lags * channels, dy).astype(np.float32)
  design = ora.add_context(x, pre, post)
  y = (design @ w_true + 0.5 + noise * rng.randn(n_frames, dy)).astype(np.float32)
  zeros = np.zeros((n_frames, 1), np.float32)
  rows = ora.file_rows(x, zeros, y, zeros, pre, post)
  batches = ora.batches_from_files([rows], batch)
  out = dict(x=x, y=y, pre=pre, post=post, batch=batch)
  configs = [('ridge0', 0.0, True, True), ('ridge1em2', 1e-2, True, True),
             ('ridge10', 10.0, True, True), ('nooffset', 0.1, False, True),
             ('shrink03', 0.3, True, False), ('lw', -1, True, False)]
  for tag, lamb, use_offset, use_ridge in configs:
    w, b, cov_x, cov_xy, shrink = (
        brain_model.calculate_linear_regressor_parameters_from_dataset(
            _dataset(batches), lamb=lamb, use_offset=use_offset, use_ridge=use_ridge))
    out[tag + '_w'] = w
    out[tag + '_b'] = np.asarray(b)
    out[tag + '_cov_x'] = cov_x
    out[tag + '_cov_xy'] = cov_xy
    out[tag + '_shrinkage'] = np.asarray(shrink, dtype=np.float64)
  np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
  print('wrote', name, {k: np.asarray(v).shape for k, v in out.items() if k.endswith('_w')})


def cca_case(name, seed, n_frames, c1, c2, post1, post2, dim, reg, batch):
  rng = np.random.RandomState(seed)
  shared = rng.randn(n_frames + 8, 3).astype(np.float32)
  x = (shared[4:4 + n_frames] @ rng.randn(3, c1) + 0.7 * rng.randn(n_frames, c1)).astype(np.float32)
  x2 = (shared[2:2 + n_frames] @ rng.randn(3, c2) + 0.7 * rng.randn(n_frames, c2)).astype(np.float32)
  y = np.ones((n_frames, 1), np.float32)
  a = np.zeros((n_frames, 1), np.float32)
  rows = ora.file_rows(x, x2, y, a, 0, post1, 0, post2)
  batches = ora.batches_from_files([rows], batch)
  rot_x, rot_y, mean_x, mean_y, e = cca.calculate_cca_parameters_from_dataset(
      _dataset(batches), dim, regularization=reg, mini_batch_count=0)
  np.savez_compressed(os.path.join(OUT, name + '.npz'), x=x, x2=x2, post1=post1,
                      post2=post2, dim=dim, reg=reg, batch=batch, rot_x=rot_x,
                      rot_y=rot_y, mean_x=mean_x, mean_y=mean_y, e=e)
  print('wrote', name, rot_x.shape, rot_y.shape, e)


def cca_test42():
  """The seeded construction of test/cca_test.py:80-123, through the reference."""
  np.random.seed(42)
  n, dim = 2000, 6
  inner = np.random.randn(n, 2)
  x = np.concatenate((inner, np.random.randn(n, dim - 2)), axis=1).astype(np.float32)
  y = np.concatenate((np.random.randn(n, dim - 2), inner @ np.array([[1., .5], [-.5, 1.]])),
                     axis=1).astype(np.float32)
  batches = [({'input_1': x[i:i + 100], 'input_2': y[i:i + 100]}, np.ones((100, 1), np.float32))
             for i in range(0, n, 100)]
  rot_x, rot_y, mean_x, mean_y, e = cca.calculate_cca_parameters_from_dataset(
      _dataset(batches), 3, regularization=0.1, mini_batch_count=0)
  np.savez_compressed(os.path.join(OUT, 'cca_seed42.npz'), x=x, y=y, rot_x=rot_x,
                      rot_y=rot_y, mean_x=mean_x, mean_y=mean_y, e=e)
  print('wrote cca_seed42', e)


class _FakeOut(object):
  def __init__(self, a):
    self._a = a
    self.shape = a.shape

  def numpy(self):
    return self._a

  def __getitem__(self, i):
    return _FakeOut(self._a[i])


def decoder_case():
  """Reference Decoder correlator + reductions + TwoResultStore windows + WTA/step."""
  rng = np.random.RandomState(7)
  n, dims, batch = 2400, 4, 200
  truth = rng.randn(n, dims).astype(np.float32)
  pred = (0.6 * truth + 0.8 * rng.randn(n, dims)).astype(np.float32)
  pred2 = rng.randn(n, dims).astype(np.float32)
  label = (np.arange(n) // 600 % 2).astype(np.float32).reshape(-1, 1)
  out = dict(truth=truth, pred=pred, pred2=pred2, label=label, batch=batch)

  class _Dec(infer_decoder.Decoder):
    def decode_one(self, input_dict, ground_truth):
      return ground_truth.numpy(), input_dict['pred'].numpy()

  for which, p in (('s1', pred), ('s2', pred2)):
    dec = _Dec(lambda d: None, reduction='all')
    for s in range(0, n, batch):
      dec.add_data_correlator(truth[s:s + batch], p[s:s + batch])
    cp = dec.correlation_params
    out[which + '_mean_x'] = np.asarray(cp.mean_x)
    out[which + '_mean_y'] = np.asarray(cp.mean_y)
    out[which + '_power'] = np.asarray(cp.power)
    out[which + '_corr'] = dec.compute_correlation(truth, p)
    for red in ('first', 'second', 'mean', 'mean-squared'):
      dec._reduction = red    # same statistics, different reducer
      batches = [({'pred': _FakeOut(p[s:s + batch]),
                   'attended_speaker': label[s:s + batch]}, _FakeOut(truth[s:s + batch]))
                 for s in range(0, n, batch)]
      for width in (10, 100, 200, 400, 700, 1000):
        res, lab = [], []
        for r, l in dec.test_by_window(batches, width):
          res.append(np.mean(r))
          lab.append(np.mean(l))
        out['%s_%s_w%d_res' % (which, red, width)] = np.asarray(res)
        out['%s_%s_w%d_lab' % (which, red, width)] = np.asarray(lab)
  # Decisions from the reference decoders on the 'first' reduction.
  for width in (10, 100, 200, 400, 700, 1000):
    r1 = out['s1_first_w%d_res' % width]
    r2 = out['s2_first_w%d_res' % width]
    wta = attention_decoder.AttentionDecoder()
    step = attention_decoder.StepAttentionDecoder()
    out['wta_w%d' % width] = np.array([wta.attention(a, b)[0] for a, b in zip(r1, r2)])
    out['step_w%d' % width] = np.array([step.attention(a, b)[0] for a, b in zip(r1, r2)])
  out['average_data_w7'] = infer_decoder.average_data(out['s1_corr'], 7)
  out['dprime'] = np.asarray(infer_decoder.calculate_dprime(out['s2_corr'][:, 0],
                                                           out['s1_corr'][:, 0]))
  # Raw window contents from TwoResultStore (result_store.py:313-338).
  store = result_store.TwoResultStore(window_width=201, window_step=100)
  sig = np.arange(1000, dtype=np.float64).reshape(-1, 1)
  firsts = []
  for s in range(0, 1000, 130):
    store.add_data(sig[s:s + 130], sig[s:s + 130] + 0.5)
    for p1, p2 in store.next_window():
      firsts.append((p1[0, 0], p1[-1, 0], p2[0, 0], p1.shape[0]))
  out['store_201_100'] = np.asarray(firsts)
  np.savez_compressed(os.path.join(OUT, 'decoder.npz'), **out)
  print('wrote decoder', len(out), 'arrays; windows@201/100:', len(firsts))


def lda_case():
  """Reference scaled LDA on the decoder fixture's correlation signals: class 0 = unattended
  speaker (s2), class 1 = attended (s1), averaged over 7-frame windows like Decoder.train."""
  g = np.load(os.path.join(OUT, 'decoder.npz'))
  c0, c1 = g['s2_corr'].astype(np.float64), g['s1_corr'].astype(np.float64)
  dec = infer_decoder.Decoder(lambda d: None, reduction='lda')
  a0, a1 = infer_decoder.average_data(c0, 7), infer_decoder.average_data(c1, 7)
  dprime = dec.compute_lda_model(a0, a1)
  params = dec.lda_params
  out = dict(dprime=np.asarray(dprime), w_real=np.asarray(params.w_real), w_imag=np.asarray(params.w_imag),
             slope=np.asarray(params.slope), intercept=np.asarray(params.intercept),
             mean_vectors=np.asarray(params.mean_vectors), avg0=a0, avg1=a1,
             transform_s1=dec.reduce_with_lda(g['s1_corr']))
  label = g['label']
  sig = dec.reduce_with_lda(g['s1_corr'])[:, 0]
  for width in (10, 100, 200, 400, 700, 1000):
    store = result_store.TwoResultStore(window_width=width, window_step=width // 2)
    res = []
    for s in range(0, sig.shape[0], 200):
      store.add_data(sig[s:s + 200], label[s:s + 200])
      for r, _ in store.next_window():
        res.append(np.mean(r))
    out['lda_w%d_res' % width] = np.asarray(res)
  lda = scaled_lda.ScaledLinearDiscriminantAnalysis()
  rng = np.random.RandomState(5)
  x = np.concatenate((rng.randn(200, 3) + [1, 2, 3], rng.randn(150, 3) + [1.5, 1, 3.5]))
  y = np.concatenate((np.full(200, 4.0), np.full(150, 9.0)))
  out['toy_x'], out['toy_y'] = x, y
  out['toy_transform'] = lda.fit_transform(x, y)
  out['toy_slope'], out['toy_intercept'] = np.asarray(lda.model_parameters.slope), np.asarray(lda.model_parameters.intercept)
  out['toy_evr'] = lda.explained_variance_ratio()
  np.savez_compressed(os.path.join(OUT, 'lda.npz'), **out)
  print('wrote lda: dprime', dprime, 'slope', params.slope)


def streaming_context_case():
  """Reference Preprocessor.add_context (preprocess.py:487-522) fed batch by batch."""
  rng = np.random.RandomState(31)
  x = rng.randn(57, 5).astype(np.float32)
  out = {'x': x}
  cases = {'p2q3': (2, 3, [0, 10, 20, 35, 57]), 'p0q4': (0, 4, [0, 6, 30, 57]),
           'p3q0': (3, 0, [0, 20, 57])}
  for tag, (pre, post, cuts) in cases.items():
    p = preprocess.Preprocessor('test', 100, 100, pre_context=pre, post_context=post)
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
      out['%s_%d' % (tag, i)] = p.add_context(x[a:b].astype(np.float64))
    out[tag + '_cuts'] = np.array(cuts)
    out[tag + '_ctx'] = np.array([pre, post])
  np.savez_compressed(os.path.join(OUT, 'streaming_context.npz'), **out)
  print('wrote streaming_context', {k: v.shape for k, v in out.items() if k.startswith('p2q3_')})


def ssd_case():
  """Reference StateSpaceAttentionDecoder (attention_decoder.py:176-451) on a stream of correlation
  pairs with an attention switch every 25 pairs: with and without the offset, with a forward lag."""
  rng = np.random.RandomState(3)
  n = 90
  att = (np.arange(n) // 25) % 2 == 0
  r1 = np.where(att, 0.25, 0.08) + 0.05 * rng.randn(n)
  r2 = np.where(att, 0.08, 0.25) + 0.05 * rng.randn(n)
  out = dict(r1=r1, r2=r2)
  settings = {'default': dict(), 'offset1': dict(offset=1.0),
              'lag3': dict(forward_lag=3, backward_lag=6)}
  for tag, kw in settings.items():
    dec = attention_decoder.StateSpaceAttentionDecoder(20, 1, 10, 1.0, **kw)
    dec.tune(np.abs(r1[:20]), np.abs(r2[:20]))
    out[tag + '_tuned'] = np.array(dec.mu_d + dec.rho_d)
    out[tag + '_out'] = np.array([dec.attention(a, b) for a, b in zip(r1, r2)], dtype=np.float64)
    out[tag + '_final'] = np.array(dec.mu_d + dec.rho_d)
  np.savez_compressed(os.path.join(OUT, 'ssd.npz'), **out)
  print('wrote ssd', out['default_out'][-1])


def main():
  os.makedirs(OUT, exist_ok=True)
  ridge_case('ridge_small', seed=11, n_frames=1000, channels=3, pre=1, post=2, dy=2,
             batch=100, noise=0.1)
  ridge_case('ridge_c8l8', seed=12, n_frames=2500, channels=8, pre=0, post=7, dy=1,
             batch=300, noise=0.3)
  cca_case('cca_small', seed=21, n_frames=3000, c1=6, c2=4, post1=3, post2=2, dim=3,
           reg=0.1, batch=250)
  cca_test42()
  decoder_case()
  lda_case()
  streaming_context_case()
  ssd_case()


if __name__ == '__main__':
  main()
