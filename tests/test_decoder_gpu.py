"""GPU parity of the decision layer: global statistics, per-frame correlation + reductions,
sliding windows for several sizes in one pass, winner-take-all decisions (bit-identical)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_numpy as ora
from telluride_decoding_b200 import attention_decoder
from telluride_decoding_b200 import brain_model
from telluride_decoding_b200 import infer
from telluride_decoding_b200 import infer_decoder
from telluride_decoding_b200 import ops
from tests import gpu_util

pytestmark = pytest.mark.gpu

WINDOWS = [10, 100, 200, 400, 700, 1000]      # infer.py:376


def _golden(golden_dir):
  return np.load(os.path.join(golden_dir, 'decoder.npz'))


def _stats(truth, pred):
  cor = ora.Correlator()
  for s in range(0, truth.shape[0], 200):
    cor.add(truth[s:s + 200], pred[s:s + 200])
  return cor


@pytest.mark.parametrize('reduction', ['first', 'second', 'mean', 'mean-squared'])
def test_windowed_global_correlation_matches_reference(golden_dir, reduction):
  g = _golden(golden_dir)
  truth, pred = g['truth'], g['pred']
  cor = _stats(truth, pred)
  t = lambda v: torch.as_tensor(np.ascontiguousarray(v), device='cuda')
  vals, off = ops.windowed_corr(t(truth), t(pred), [0], [truth.shape[0]], WINDOWS,
                                [w // 2 for w in WINDOWS], 'global', reduction=reduction,
                                mu_a=t(cor.mean_x), mu_b=t(cor.mean_y), power=t(cor.power))
  vals = vals.cpu().numpy()[:, 0]
  for i, w in enumerate(WINDOWS):
    want = g['s1_%s_w%d_res' % (reduction, w)]
    got = vals[off[i]:off[i + 1]]
    assert got.shape == want.shape                      # window count (N - W)//S + 1
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_labels_windows_and_wta_decisions_bit_identical(golden_dir):
  g = _golden(golden_dir)
  t = lambda v: torch.as_tensor(np.ascontiguousarray(v), device='cuda')
  n = g['truth'].shape[0]
  res = {}
  for which, p in (('s1', g['pred']), ('s2', g['pred2'])):
    cor = _stats(g['truth'], p)
    vals, off = ops.windowed_corr(t(g['truth']), t(p), [0], [n], WINDOWS, [w // 2 for w in WINDOWS],
                                  'global', reduction='first', mu_a=t(cor.mean_x),
                                  mu_b=t(cor.mean_y), power=t(cor.power))
    res[which] = (vals, off)
  labs, off = ops.windowed_corr(t(g['label']), None, [0], [n], WINDOWS, [w // 2 for w in WINDOWS],
                                'plain')
  labs = labs.cpu().numpy()[:, 0]
  for i, w in enumerate(WINDOWS):
    np.testing.assert_array_equal(labs[off[i]:off[i + 1]], g['s1_first_w%d_lab' % w])
    r1 = res['s1'][0][off[i]:off[i + 1], 0]
    r2 = res['s2'][0][off[i]:off[i + 1], 0]
    wta = ops.wta_decide(r1, r2).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(wta, g['wta_w%d' % w])
    np.testing.assert_array_equal(attention_decoder.step_decisions(wta), g['step_w%d' % w])


def test_column_moments_and_multi_segment_windows():
  rng = np.random.RandomState(3)
  a = rng.randn(5000, 3).astype(np.float32) + 1
  b = rng.randn(5000, 3).astype(np.float32) * 2
  t = lambda v: torch.as_tensor(v, device='cuda')
  sums = ops.column_moments(t(a), t(b)).cpu().numpy()
  a64, b64 = a.astype(np.float64), b.astype(np.float64)
  np.testing.assert_allclose(sums, [a64.sum(0), b64.sum(0), (a64 ** 2).sum(0), (b64 ** 2).sum(0)],
                             rtol=1e-12)
  # windows never straddle trials: three segments, odd window shape (gcd 1 path)
  seg_start, seg_len = [0, 1500, 3100], [1500, 1600, 1900]
  vals, off = ops.windowed_corr(t(a), None, seg_start, seg_len, [301, 64], [150, 32], 'plain')
  vals = vals.cpu().numpy()[:, 0]
  want = []
  for w, s in ((301, 150), (64, 32)):
    for st, n in zip(seg_start, seg_len):
      sig = a64[st:st + n, 0]
      want += [sig[i:i + w].mean() for i in ora.window_starts(n, w, s)]
  np.testing.assert_allclose(vals, want, rtol=1e-12)
  assert off[1] == sum(ops.window_count(n, 301, 150) for n in seg_len)
  from telluride_decoding_b200 import _lib
  with pytest.raises(_lib.TdError, match='step 0 never terminates'):
    ops.windowed_corr(t(a), None, [0], [100], [1], [0], 'plain')


def test_local_pearson_windows_and_zero_rule():
  rng = np.random.RandomState(5)
  a = rng.randn(3000, 4).astype(np.float32) + 3
  b = (0.5 * a + rng.randn(3000, 4)).astype(np.float32)
  t = lambda v: torch.as_tensor(v, device='cuda')
  vals, _ = ops.windowed_corr(t(a), t(b), [0], [3000], [500], [500], 'local')
  want = np.array([ora.pearson_correlation(a[i:i + 500].astype(np.float64),
                                           b[i:i + 500].astype(np.float64))
                   for i in range(0, 3000, 500)])
  np.testing.assert_allclose(vals.cpu().numpy(), want, rtol=1e-9, atol=1e-12)
  a[:, 2] = 7.0      # a constant column zeroes the whole vector (brain_model.py:72-79)
  vals, _ = ops.windowed_corr(t(a), t(b), [0], [3000], [500], [500], 'local')
  assert np.all(vals.cpu().numpy() == 0)


def test_decoder_end_to_end_attention_accuracy():
  """Two-speaker synthetic experiment in the spirit of test/infer_test.py:158-176: the attended
  envelope drives the EEG; accuracy > 0.95 for windows >= 100 frames, and the window values equal
  the oracle's run on the same model outputs."""
  rng = np.random.RandomState(21)
  n, c, post = 12000, 16, 7
  env1 = rng.randn(n + post, 1).astype(np.float32)
  env2 = rng.randn(n + post, 1).astype(np.float32)
  label = (np.arange(n) >= n // 2).astype(np.float32).reshape(-1, 1)   # 0: speaker 1 attended, one switch
  att = np.where(label > 0, env2[:n], env1[:n])
  mix = rng.randn(post + 1, c).astype(np.float32)
  eeg = np.zeros((n, c), np.float32)
  for k in range(post + 1):          # EEG lags the attended envelope
    eeg[k:] += att[:n - k] * mix[k]
  eeg += 0.5 * rng.randn(n, c).astype(np.float32)
  kw = dict(post_context=post, final_batch_size=200)
  bd_fit = gpu_util.make_test_data(eeg, att.astype(np.float32), a=label, **kw)
  ds_fit = bd_fit.create_dataset('train')
  model = brain_model.BrainModelLinearRegression(ds_fit, regularization_lambda=0.1)
  model.fit(ds_fit)
  bd1 = gpu_util.make_test_data(eeg, env1[:n], a=label, **kw)
  bd2 = gpu_util.make_test_data(eeg, env2[:n], a=label, **kw)
  dec = infer_decoder.create_decoder('linear', reduction='first', model=model)
  dec.train_statistics(bd1.create_dataset('train'), bd2.create_dataset('train'))
  acc = infer.run_reduction_test(dec, bd1.create_dataset('test'), bd2.create_dataset('test'), 'wta')
  # only the one or two windows that straddle the switch can be wrong
  assert all(acc[w] > 0.9 for w in WINDOWS if w >= 100), acc
  assert acc[10] > 0.6
  acc_step = infer.run_reduction_test(dec, bd1.create_dataset('test'), bd2.create_dataset('test'),
                                      'stepped')
  assert acc_step[1000] > 0.8
  # oracle on the same predictions
  pred = model.predict(bd1.create_dataset('test'))
  cor = ora.Correlator()
  for e in (env1[:n], env2[:n]):
    for s in range(0, n, 200):
      cor.add(e[s:s + 200], pred[s:s + 200])
  assert np.allclose(np.float32(dec.correlation_params.mean_x), cor.mean_x, rtol=1e-4, atol=1e-6)
  assert np.allclose(np.float32(dec.correlation_params.power), cor.power, rtol=1e-4)
  res = dec.windowed_results(bd1.create_dataset('test'), [200])[200]
  sig = ora.reduce_correlations(cor.correlate(env1[:n], pred), 'first')
  want, want_lab = ora.windowed_means(sig, label, 200)
  np.testing.assert_allclose(res[0], want, rtol=1e-3, atol=1e-5)
  np.testing.assert_array_equal(res[1], want_lab)


def test_lda_training_and_windows_match_reference(golden_dir):
  """Decoder.train (global statistics + scaled LDA on 7-frame averages) and the 'lda' reducer in
  the window kernel against the reference's Decoder / scaled_lda run on the same signals."""
  g = _golden(golden_dir)
  lg = np.load(os.path.join(golden_dir, 'lda.npz'))
  t = lambda v: torch.as_tensor(np.ascontiguousarray(v), device='cuda')
  n = g['truth'].shape[0]

  class FixedDecoder(infer_decoder.Decoder):
    """decode_dataset returns prepared model outputs; each class uses its own statistics like the
    fixture (the reference fixture trained one correlator per speaker)."""
    def decode_dataset(self, dataset):
      return dataset

  # class-wise correlation signals with the fixture's per-speaker statistics, all columns
  avg = []
  for which, p in (('s2', g['pred2']), ('s1', g['pred'])):
    cor = _stats(g['truth'], p)
    vals, _ = ops.windowed_corr(t(g['truth']), t(p), [0], [n], [7], [7], 'global_all',
                                mu_a=t(cor.mean_x), mu_b=t(cor.mean_y), power=t(cor.power))
    avg.append(vals.cpu().numpy())
  np.testing.assert_allclose(avg[0], lg['avg0'], rtol=1e-12, atol=1e-14)
  np.testing.assert_allclose(avg[1], lg['avg1'], rtol=1e-12, atol=1e-14)
  dec = FixedDecoder(lambda d: None, reduction='lda')
  dprime = dec.compute_lda_model(avg[0], avg[1])
  assert abs(dprime - float(lg['dprime'])) < 1e-8
  # 'lda' windows of speaker 1 through the kernel
  cor = _stats(g['truth'], g['pred'])
  dec._set_correlation_params([cor.count, cor.sum_x, cor.sum_y, cor.sum_x2, cor.sum_y2,
                               cor.mean_x, cor.mean_y, cor.power])
  bd = gpu_util.make_test_data(g['truth'], g['truth'][:, :1], a=g['label'], final_batch_size=200)
  ds = bd.create_dataset('test')
  dec.decode_dataset = lambda dataset: (t(g['truth']), t(g['pred']))
  res = dec.windowed_results(ds, WINDOWS)
  for w in WINDOWS:
    np.testing.assert_allclose(res[w][0], lg['lda_w%d_res' % w], rtol=1e-9, atol=1e-11)
    np.testing.assert_array_equal(res[w][1], g['s1_first_w%d_lab' % w])
  # the whole-dataset train(): statistics pooled over both classes, then LDA; d' is large
  dec2 = FixedDecoder(lambda d: None, reduction='lda')
  dec2.decode_dataset = lambda dataset: dataset
  class _DS(tuple):
    pass
  import telluride_decoding_b200.brain_model as bm
  orig = bm.is_dataset
  bm.is_dataset = lambda o: True
  try:
    d2 = dec2.train((t(g['truth']), t(g['pred2'])), (t(g['truth']), t(g['pred'])), window_size=7)
  finally:
    bm.is_dataset = orig
  pooled = ora.Correlator()
  for p in (g['pred2'], g['pred']):
    for s in range(0, n, 200):
      pooled.add(g['truth'][s:s + 200], p[s:s + 200])
  c0 = ora.average_data(pooled.correlate(g['truth'], g['pred2']).astype(np.float64), 7)
  c1 = ora.average_data(pooled.correlate(g['truth'], g['pred']).astype(np.float64), 7)
  ref = infer_decoder.Decoder(lambda d: None, reduction='lda')
  want = ref.compute_lda_model(c0, c1)
  assert abs(d2 - want) < 1e-3 * abs(want) and d2 > 1.0


def _aad_trials(seed, lens, c, gain_unattended=0.1):
  """Two speakers per trial; the attended one drives the EEG (test/decoding_test.py:134-216 style),
  the label alternates between trials."""
  rng = np.random.RandomState(seed)
  trials = []
  taps = rng.randn(5, c).astype(np.float32) / 3
  for i, n in enumerate(lens):
    s1 = rng.randn(n, 1).astype(np.float32)
    s2 = rng.randn(n, 1).astype(np.float32)
    label = i % 2                                   # 0: speaker 1 attended
    att, un = (s1, s2) if label == 0 else (s2, s1)
    eeg = 0.8 * rng.randn(n, c).astype(np.float32)
    for k in range(5):                              # the EEG lags the audio: decode with post context
      eeg[k:] += att[:n - k] * taps[k] + gain_unattended * un[:n - k] * taps[k]
    trials.append(dict(eeg=eeg, att=att.astype(np.float32), s1=s1, s2=s2,
                       label=np.full((n, 1), label, np.float32)))
  return trials


def _oracle_loo_aad(trials, order, pre, post, batch, lam, windows, reduction):
  """The reference's loop structure run with the oracle: one complete fit per held-out trial
  (regression.py:197-214), held-out prediction, pooled correlator (infer_decoder.py:288-328),
  windows inside each trial (result_store.py:262-271), winner-take-all, accuracy (infer.py:399-407)."""
  rows = {}
  for name in order:
    t = trials[name]
    z = np.zeros((t['eeg'].shape[0], 1), np.float32)
    rows[name] = ora.file_rows(t['eeg'], z, t['att'], z, pre, post)
  preds = {}
  for held in order:
    batches = ora.batches_from_files([rows[n] for n in order if n != held], batch)
    w, b, _, _, _ = ora.linear_regressor_from_batches(batches, lam)
    preds[held] = (np.matmul(rows[held][0], w) + np.reshape(b, (1, -1))).astype(np.float32)
  corr = ora.Correlator()
  pred_all = np.concatenate([preds[n] for n in order])
  corr.add(np.concatenate([trials[n]['s1'] for n in order]), pred_all)
  corr.add(np.concatenate([trials[n]['s2'] for n in order]), pred_all)
  out = {}
  for w_ in windows:
    r1, r2, lab = [], [], []
    for n in order:
      t = trials[n]
      c1 = ora.reduce_correlations(corr.correlate(t['s1'], preds[n]), reduction)
      c2 = ora.reduce_correlations(corr.correlate(t['s2'], preds[n]), reduction)
      a, l = ora.windowed_means(c1, t['label'], w_)
      b2, _ = ora.windowed_means(c2, t['label'], w_)
      r1.append(a); r2.append(b2); lab.append(l)
    r1, r2, lab = np.concatenate(r1), np.concatenate(r2), np.concatenate(lab)
    att = ora.wta_attention(r1, r2)
    out[w_] = dict(r1=r1, r2=r2, labels=lab, attention=att, accuracy=ora.aad_accuracy(att, lab))
  return out, preds


@pytest.mark.parametrize('lens,batch', [
    ([1500] * 20, 100),                               # 20 trials, no row is dropped in any fold
    ([1300, 900, 1111, 1200, 777, 1024], 128),        # ragged trials: every fold drops a different tail
])
def test_leave_one_trial_out_aad_matches_the_reference_loop(lens, batch):
  """BASELINE.json configs[4] end to end on a subset the oracle can loop over: per-trial statistics
  from one launch, batched fold solves, held-out predictions, correlator, windows, decisions --
  decisions bit-identical, correlations within 1e-4."""
  from telluride_decoding_b200 import brain_data, infer
  c, pre, post, lam = 8, 0, 7, 0.1
  trials_l = _aad_trials(11, lens, c)
  names = ['trial_%02d' % i for i in range(len(lens))]
  trials = dict(zip(names, trials_l))
  bd = brain_data.TrialBrainData('eeg', 'attended', 100.0, pre_context=pre, post_context=post,
                                 in2_fields='speakers', attended_field='label',
                                 final_batch_size=batch, shuffle_buffer_size=0)
  for n in names:
    t = trials[n]
    bd.add_trial(n, t['eeg'], t['att'], np.concatenate([t['s1'], t['s2']], axis=1), t['label'])
  windows = [50, 100, 200]
  acc, det = infer.leave_one_trial_out_test(bd, regularization_lambda=lam, window_list=windows,
                                            reduction='first', decoder_type='wta', return_details=True)
  order = bd.all_files()
  want, preds = _oracle_loo_aad(trials, order, pre, post, batch, lam, windows, 'first')
  pred_want = np.concatenate([preds[n] for n in order])
  assert gpu_util.rel_fro(det['prediction'].cpu().numpy(), pred_want) < 1e-4
  for w_ in windows:
    np.testing.assert_allclose(det[w_]['r1'], want[w_]['r1'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(det[w_]['r2'], want[w_]['r2'], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(det[w_]['labels'], want[w_]['labels'])
    np.testing.assert_array_equal(det[w_]['attention'], want[w_]['attention'])      # bit-identical decisions
    assert acc[w_] == want[w_]['accuracy']
  assert acc[200] > 0.8                                # the attended speaker is decodable
  # equal trials and no dropped rows: one factorisation + fixed-point steps; otherwise the batched Cholesky
  assert det['solver'].startswith('one factorisation') == (batch == 100), det['solver']
