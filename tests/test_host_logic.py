"""CPU: host-side logic of the package against the oracle / the reference's goldens --
row plans (interleave, drop_remainder, chunks, offsets), stores, decoders, options."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_numpy as ora
from telluride_decoding_b200 import attention_decoder
from telluride_decoding_b200 import brain_data
from telluride_decoding_b200 import decoding
from telluride_decoding_b200 import infer
from telluride_decoding_b200 import infer_decoder
from telluride_decoding_b200 import regression
from telluride_decoding_b200 import result_store


@pytest.mark.parametrize('lengths,batch', [([3, 1, 2], 4), ([10, 10, 10], 7), ([5, 50, 20, 1], 16),
                                           ([100], 33), ([7, 0, 9], 5)])
def test_kept_prefix_matches_oracle_interleave(lengths, batch):
  fi, _ = ora.interleave_order(lengths)
  kept_total = (len(fi) // batch) * batch
  want = [int(np.sum(fi[:kept_total] == f)) for f in range(len(lengths))]
  assert brain_data.interleave_kept_prefix(lengths, kept_total) == want
  f2, r2 = brain_data.interleave_order(lengths)
  np.testing.assert_array_equal(f2, fi)
  np.testing.assert_array_equal(r2, ora.interleave_order(lengths)[1])


def test_kept_prefix_every_cut_of_random_file_sets():
  """Every cut position of 200 random file sets against the oracle's explicit stream order."""
  import random
  rng = random.Random(0)
  for _ in range(200):
    lengths = [rng.randint(0, 12) for _ in range(rng.randint(1, 7))]
    fi, _ = ora.interleave_order(lengths)
    for cut in range(0, len(fi) + 2):
      want = [int(np.sum(fi[:cut] == f)) for f in range(len(lengths))]
      assert brain_data.interleave_kept_prefix(lengths, cut) == want, (lengths, cut)


def _dataset(files_len, pre=0, post=0, pre2=0, post2=0, offset=0, chunk=1000000, batch=10,
             mode='test', repeat=1, offset_counts=True):
  n = sum(files_len)
  x = torch.zeros((n, 3))
  x2 = torch.zeros((n, 2))
  y = torch.zeros((n, 1))
  a = torch.zeros((n, 1))
  files, cur = [], 0
  for i, l in enumerate(files_len):
    files.append(('f%d' % i, cur, l))
    cur += l
  return brain_data.LaggedDataset(x, x2, y, a, files, pre, post, pre2, post2, offset, chunk, batch,
                                  repeat, 0, mode, offset_counts=offset_counts)


def test_segments_chunks_and_offsets_follow_window_data():
  ds = _dataset([25], pre=1, post=1, offset=2, chunk=10, batch=4)
  # chunks of 10, 10, 5 frames; offset 2 drops 2 rows of x per chunk (brain_data.py:466-470)
  assert ds.seg_rows == [8, 8, 3]
  sx = ds.stream('x')
  assert sx.seg_start == [2, 12, 22] and sx.seg_len == [8, 8, 3]
  sy = ds.stream('y')
  assert sy.seg_start == [0, 10, 20] and sy.seg_len == [10, 10, 5]
  assert ds.total_rows == 19 and ds.num_batches == 4
  assert ds.seg_keep_for_prefix(16) == [8, 8, 0]
  ds = _dataset([25], pre=1, offset=-3, chunk=10, batch=4)
  assert ds.seg_rows == [7, 7, 2]
  assert ds.stream('y').seg_start == [3, 13, 23]
  assert ds.stream('x').seg_start == [0, 10, 20] and ds.stream('x').seg_len == [10, 10, 5]
  # no context at all: one segment per file, offset ignored for TFExampleData-like sources
  ds = _dataset([25], offset=2, offset_counts=False)
  assert ds.seg_rows == [25] and ds.input_offset == 0
  assert ds.element_spec[0]['input_1'].shape[-1] == 3


def test_passes_cover_repeat_and_remainder():
  ds = _dataset([10, 4], batch=4, mode='train', repeat=3)      # 42 rows -> 10 batches = 40 rows
  assert ds.num_batches == 10
  (w1, k1), (w2, k2) = ds.passes()
  assert w1 == 2.0 and k1 == [10, 4]
  assert w2 == 1.0 and sum(k2) == 12
  assert k2 == brain_data.interleave_kept_prefix([10, 4], 12)
  ds = _dataset([10, 4], batch=7, mode='test', repeat=3)      # repeat only applies in train mode
  assert ds.num_batches == 2 and ds.passes() == [(1.0, [10, 4])]


def test_file_filters_allbut(monkeypatch):
  monkeypatch.setattr(brain_data._lib, 'as_device_f32',
                      lambda a, device=None: torch.as_tensor(np.asarray(a, dtype=np.float32)))
  bd = brain_data.TrialBrainData('eeg', 'intensity', 100.0, final_batch_size=5,
                                 shuffle_buffer_size=0)
  for i in range(4):
    bd.add_trial('trial_%02d' % i, np.zeros((20 + i, 2)), np.zeros((20 + i, 1)))
  assert sorted(bd.all_files()) == ['trial_00', 'trial_01', 'trial_02', 'trial_03']
  bd.set_file_patterns('allbut', 'trial_02', 'trial_02')
  assert sorted(bd.filter_file_names('train')) == ['trial_00', 'trial_01', 'trial_03']
  assert bd.filter_file_names('test') == ['trial_02']
  ds = bd.create_dataset('train')
  assert ds.total_rows == 20 + 21 + 23 and ds.num_batches == 12
  bd.set_file_patterns('allbut', '', 'trial_02')
  with pytest.raises(ValueError, match='Both test and validate must be specified'):
    bd.filter_file_names('train')
  with pytest.raises(ValueError, match='mode must be one of'):
    bd.filter_file_names('bogus')
  assert bd.input_fields_width() == 2 and bd.output_field_width() == 1


def test_brain_data_argument_errors():
  with pytest.raises(ValueError, match='pre_context must be >= 0'):
    brain_data.TestBrainData('in', 'out', 100.0, pre_context=-1)
  with pytest.raises(ValueError, match='Must specify at least one input field'):
    brain_data.TestBrainData('', 'out', 100.0)
  with pytest.raises(ValueError, match='Must call preserve_test_data'):
    brain_data.TestBrainData('in', 'out', 100.0).create_dataset('train')
  with pytest.raises(TypeError, match='unknown data type'):
    brain_data.create_brain_dataset('nope', 'in', 'out', 100.0)
  with pytest.raises(ValueError, match='frame_rate must be greater than 0'):
    brain_data.create_brain_dataset('test', 'in', 'out', 0)


def test_result_store_windows_match_reference(golden_dir):
  g = np.load(os.path.join(golden_dir, 'decoder.npz'))
  store = result_store.TwoResultStore(window_width=201, window_step=100)
  sig = np.arange(1000, dtype=np.float64).reshape(-1, 1)
  got = []
  for s in range(0, 1000, 130):
    store.add_data(sig[s:s + 130], sig[s:s + 130] + 0.5)
    for p1, p2 in store.next_window():
      got.append((p1[0, 0], p1[-1, 0], p2[0, 0], p1.shape[0]))
  np.testing.assert_array_equal(np.asarray(got), g['store_201_100'])
  np.testing.assert_array_equal(result_store.window_starts(1000, 201, 100), np.arange(8) * 100)
  # pre_context primes store 1 with zeros (test/result_store_test.py:112-147)
  ws = result_store.WindowedDataStore(window_step=2, window_width=4, pre_context=2)
  ws.add_data(np.arange(1, 9, dtype=float).reshape(-1, 1))
  wins = [w[:, 0].tolist() for w in ws.next_window()]
  assert wins == [[0, 0, 1, 2], [1, 2, 3, 4], [3, 4, 5, 6], [5, 6, 7, 8]]
  ns = result_store.NumpyStore(init_frame_count=4)
  for i in range(5):
    ns.add_data(np.full((3, 2), i, dtype=float))
  assert ns.count == 15 and ns.all_data.shape == (15, 2) and ns.all_data[-1, 0] == 4
  first = next(ns.next_window(6))
  assert first.shape == (6, 2) and ns.count == 9
  with pytest.raises(ValueError, match='window_step'):
    result_store.WindowedDataStore(window_step=5, window_width=4)


def test_streaming_decoder_matches_reference(golden_dir):
  g = np.load(os.path.join(golden_dir, 'decoder.npz'))
  truth, pred, batch = g['truth'], g['pred'], int(g['batch'])

  class Dec(infer_decoder.Decoder):
    def decode_one(self, input_dict, ground_truth):
      return ground_truth, input_dict['pred']

  dec = Dec(lambda d: None, reduction='mean-squared')
  for s in range(0, truth.shape[0], batch):
    dec.add_data_correlator(truth[s:s + batch], pred[s:s + batch])
  np.testing.assert_array_equal(dec.correlation_params.power, g['s1_power'])
  np.testing.assert_array_equal(dec.compute_correlation(truth, pred), g['s1_corr'])
  batches = [({'pred': pred[s:s + batch], 'attended_speaker': g['label'][s:s + batch]},
              truth[s:s + batch]) for s in range(0, truth.shape[0], batch)]
  res = [np.mean(r) for r, _ in dec.test_by_window(batches, 200)]
  np.testing.assert_allclose(res, g['s1_mean-squared_w200_res'], rtol=0, atol=1e-15)
  with pytest.raises(ValueError, match='Unknown reduction technique'):
    infer_decoder.Decoder(reduction='bogus')
  with pytest.raises(TypeError, match='callable model'):
    infer_decoder.Decoder(decoding_model=3)
  np.testing.assert_array_equal(
      infer_decoder.average_data(np.arange(12, dtype=float).reshape(6, 2), 2),
      [[1, 2], [5, 6], [9, 10]])
  assert abs(infer_decoder.calculate_dprime(g['s2_corr'][:, 0], g['s1_corr'][:, 0]) -
             float(g['dprime'])) < 1e-12


def test_attention_decoders_reference_sequences():
  cor2 = [1] * 13
  cor1 = [2] * 7 + [0] * 6
  ad = attention_decoder.create_attention_decoder('wta')
  assert [ad.attention(a, b)[0] for a, b in zip(cor1, cor2)] == [True] * 7 + [False] * 6
  sd = attention_decoder.create_attention_decoder('stepped')
  assert [sd.attention(a, b)[0] for a, b in zip(cor1, cor2)] == [True] * 11 + [False] * 2
  np.testing.assert_array_equal(attention_decoder.step_decisions([True] * 2 + [False] * 11),
                                [True] * 3 + [False] * 10)
  with pytest.raises(ValueError, match='Unknown type'):
    attention_decoder.create_attention_decoder('bogus')
  assert infer.find_first_segment(np.array([0, 0, 0, 1, 1, 0])) == 3
  assert infer.find_first_segment(np.array([1, 1, 1])) == 0


def test_regularization_parsing_and_options():
  np.testing.assert_allclose(regression.parse_regularization_values('normal'),
                             10.0 ** np.arange(-6, 1))
  np.testing.assert_allclose(regression.parse_regularization_values('0.1,1,10'), [0.1, 1, 10])
  assert regression.parse_regularization_values(0.5) == [0.5]
  with pytest.raises(TypeError):
    regression.parse_regularization_values(3)
  opts = decoding.DecodingOptions()
  assert opts.batch_size == 512 and opts.test_metric == 'pearson_correlation_first'
  r = regression.RegressionLinear(opts)
  assert r.preset_flags() == 'linear'
  assert opts.post_context == 20 and opts.batch_size == 100 and opts.shuffle_buffer_size == 0
  r = regression.RegressionCCA(decoding.DecodingOptions())
  r.preset_flags()
  assert (r.my_flags.post_context, r.my_flags.input2_pre_context, r.my_flags.cca_dimensions) == (21, 15, 5)
  assert 'batch_size=100' in opts.experiment_parameters()
  with pytest.raises(TypeError, match='DecodingOptions'):
    decoding.create_brain_model({}, None)


def test_scaled_lda_matches_reference(golden_dir):
  """Host-side LDA (float64 numpy) against the unmodified reference's scaled_lda / Decoder."""
  from telluride_decoding_b200 import scaled_lda
  g = np.load(os.path.join(golden_dir, 'lda.npz'))
  lda = scaled_lda.ScaledLinearDiscriminantAnalysis()
  got = lda.fit_transform(g['toy_x'], g['toy_y'])
  # two classes: inv(Sw) Sb has rank one, so only the first axis is determined (the second is an
  # arbitrary null-space vector that depends on summation order)
  np.testing.assert_allclose(got[:, 0], g['toy_transform'][:, 0], rtol=1e-9, atol=1e-10)
  np.testing.assert_allclose(lda.model_parameters.slope, g['toy_slope'], rtol=1e-9)
  np.testing.assert_allclose(lda.explained_variance_ratio()[0], g['toy_evr'][0], rtol=1e-7)
  assert lda.labels == [4.0, 9.0]
  with pytest.raises(ValueError, match='two-class data'):
    scaled_lda.ScaledLinearDiscriminantAnalysis().fit(g['toy_x'], np.arange(350) % 3)
  with pytest.raises(ValueError, match='Must fit the model'):
    scaled_lda.LinearDiscriminantAnalysis().transform(g['toy_x'])
  # Decoder.compute_lda_model / reduce_with_lda on the averaged correlation signals
  dec = infer_decoder.Decoder(lambda d: None, reduction='lda')
  dprime = dec.compute_lda_model(g['avg0'], g['avg1'])
  assert abs(dprime - float(g['dprime'])) < 1e-9
  p = dec.lda_params
  np.testing.assert_allclose(p.w_real[:, 0], g['w_real'][:, 0], rtol=1e-8, atol=1e-12)
  np.testing.assert_allclose(p.slope, g['slope'], rtol=1e-9)
  d = np.load(os.path.join(golden_dir, 'decoder.npz'))
  np.testing.assert_allclose(dec.reduce_with_lda(d['s1_corr'])[:, 0], g['transform_s1'][:, 0],
                             rtol=1e-9, atol=1e-12)
  with pytest.raises(ValueError, match='Must compute the LDA model'):
    infer_decoder.Decoder(lambda d: None, reduction='lda').reduce_with_lda(d['s1_corr'])
  # parameters survive a round trip through model_params
  other = infer_decoder.Decoder(lambda d: None, reduction='lda')
  other.model_params = dec.model_params
  np.testing.assert_allclose(other.reduce_with_lda(d['s1_corr'])[:, 0], g['transform_s1'][:, 0],
                             rtol=1e-9, atol=1e-12)


def test_total_windows_matches_the_store_rule():
  """(n - W)//S + 1 windows per segment with n >= W (test/result_store_test.py:217-218)."""
  from telluride_decoding_b200 import ops
  lens = [1000, 7680, 99, 128, 0, 3841]
  widths, steps = [128, 256, 640, 3840, 201], [64, 128, 320, 1920, 100]
  want = sum((n - w) // s + 1 for w, s in zip(widths, steps) for n in lens if n >= w)
  assert ops.total_windows(lens, widths, steps) == want
  assert ops.total_windows([1000], [201], [100]) == 8
  assert ops.total_windows([], [10], [5]) == 0


def test_upload_chunks_and_segment_groups():
  """Trial-aligned upload chunks (10 / 40 / 100 % of the frames) and the grouping of dataset
  segments by the chunk that completes them (events replaced by tags: no GPU needed)."""
  import torch
  from telluride_decoding_b200 import brain_data
  assert brain_data.TrialBrainData.upload_cuts([7680] * 60) == [46080, 184320, 460800]
  assert brain_data.TrialBrainData.upload_cuts([700, 650, 900, 512, 801, 640, 777]) == [700, 2250, 4980]
  assert brain_data.TrialBrainData.upload_cuts([5, 5]) == [5, 10]
  assert brain_data.TrialBrainData.upload_cuts([10, 1, 1]) == [10, 11, 12]
  lengths = [700, 650, 900, 512, 801, 640, 777]
  n = sum(lengths)
  x = torch.zeros((n, 3)); y = torch.zeros((n, 1)); z = torch.zeros((n, 1))
  files, cur = [], 0
  order = [3, 0, 6, 1, 5, 2, 4]                      # dataset file order differs from the row order
  starts = np.concatenate(([0], np.cumsum(lengths)[:-1]))
  for i in order:
    files.append(('t%d' % i, int(starts[i]), lengths[i]))
  ds = brain_data.LaggedDataset(x, z, y, z, files, 1, 2, 0, 0, 0, 1000000, 100, 1, 0, 'train')
  assert ds.segment_groups() == [(None, list(range(7)))]
  ds.ready = [(700, 'ev0'), (2250, 'ev1'), (4980, 'ev2')]
  groups = ds.segment_groups()
  assert [ev for ev, _ in groups] == ['ev0', 'ev1', 'ev2']
  by_event = {ev: sorted(order[s] for s in segs) for ev, segs in groups}
  assert by_event == {'ev0': [0], 'ev1': [1, 2], 'ev2': [3, 4, 5, 6]}
  assert sorted(s for _, segs in groups for s in segs) == list(range(7))


def test_state_space_decoder_matches_reference(golden_dir):
  """StateSpaceAttentionDecoder against the unmodified reference class (tests/golden/ssd.npz from
  oracle/gen_golden.py): every output triple of a 90-pair stream, three settings."""
  from telluride_decoding_b200 import attention_decoder
  g = np.load(os.path.join(golden_dir, 'ssd.npz'))
  r1, r2 = g['r1'], g['r2']
  settings = {'default': dict(), 'offset1': dict(offset=1.0),
              'lag3': dict(forward_lag=3, backward_lag=6)}
  for tag, kw in settings.items():
    dec = attention_decoder.StateSpaceAttentionDecoder(20, 1, 10, 1.0, **kw)
    dec.tune(np.abs(r1[:20]), np.abs(r2[:20]))
    np.testing.assert_allclose(np.array(dec.mu_d + dec.rho_d), g[tag + '_tuned'], rtol=1e-13)
    got = np.array([dec.attention(a, b) for a, b in zip(r1, r2)], dtype=np.float64)
    np.testing.assert_allclose(got, g[tag + '_out'], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(np.array(dec.mu_d + dec.rho_d), g[tag + '_final'], rtol=1e-11)
    assert (got[:dec.k_w - 1] == 0.5).all()            # undecided until the window is full


def test_state_space_decoder_like_the_reference_tests():
  """test/attention_decoder_test.py:184-236 (decodes synthetic switches, error < 15 %), :238-284
  (log-normal initialisation) and :291-300 (factory)."""
  from telluride_decoding_b200 import attention_decoder
  rng = np.random.RandomState(0)
  n, switch, init = 1000, 300, 100
  true_state = ((np.arange(n) // switch) % 2) + 1
  att, unatt = 0.2 + 0.1 * rng.randn(n), 0.05 + 0.1 * rng.randn(n)
  corr = np.where((true_state == 1)[:, None], np.stack([att, unatt], 1), np.stack([unatt, att], 1))
  corr = np.minimum(1, np.maximum(-1, corr))
  ad = attention_decoder.StateSpaceAttentionDecoder(20, 1, 10, 1)
  ad.tune_log_normal_priors(corr[:init, 0], corr[:init, 1])
  results = np.zeros((n - ad.k_w - init, 3))
  for i in range(n - init):
    results[i - ad.k_w] = ad.attention(corr[i + init, 0], corr[i + init, 1])
  s1 = results[:, 0] > 0.5
  assert np.mean(s1 != (true_state[init + ad.k_w:] < 2)) < 0.15
  data_a = np.exp(rng.randn(200000) * 0.1 + 0.2)
  data_b = np.exp(rng.randn(200000) * 0.1)
  ad = attention_decoder.StateSpaceAttentionDecoder(20, 1, 10, 1, 1)
  ad.tune(data_a, data_b)
  assert abs(ad.mu_d[0] - np.mean(np.log(data_a))) < 1e-3
  assert abs(ad.rho_d[0] - 1.0 / np.var(np.log(data_a))) < 5
  ad2 = attention_decoder.StateSpaceAttentionDecoder(20, 1, 10, 1, 1, offset=1.0)
  ad2.tune(data_a, data_b)
  assert ad2.mu_d[0] > np.mean(np.log(data_a)) + 0.01
  assert isinstance(attention_decoder.create_attention_decoder('ssd'),
                    attention_decoder.StateSpaceAttentionDecoder)
  with pytest.raises(ValueError, match='Unknown type'):
    attention_decoder.create_attention_decoder('bad type name')


def test_batch_pieces_are_the_minibatches_of_the_interleaved_stream():
  """LaggedDataset.batch_pieces (the row plan of the Ledoit-Wolf kernels): minibatch b of the
  interleaved stream is a contiguous row range of every file, as ora.batches_from_files lays it out
  (brain_data.py:725-728, :369-370)."""
  lens = [530, 210, 377]
  c = 2
  rng = np.random.RandomState(1)
  xs = [rng.randn(n, c).astype(np.float32) for n in lens]
  files, cur = [], 0
  for i, n in enumerate(lens):
    files.append(('f%d' % i, cur, n))
    cur += n
  x = torch.from_numpy(np.concatenate(xs))
  z = torch.zeros((cur, 1))
  batch = 100
  ds = brain_data.LaggedDataset(x, z, z, z, files, 0, 0, 0, 0, 0, 1000000, batch, 1, 0, 'test',
                                temporal_context=False)
  pieces = ds.batch_pieces()
  # the oracle's stream: tag every row with (file, row) and batch the interleaved stream
  tagged = [(np.stack([np.full(n, i), np.arange(n)], axis=1).astype(np.float32),) * 4 for i, n in enumerate(lens)]
  batches = ora.batches_from_files(tagged, batch)
  assert len(batches) == ds.num_batches == sum(lens) // batch
  for b, (xd, _) in enumerate(batches):
    tags = np.asarray(xd['input_1']).astype(int)
    mine = [(s, lo, hi) for s, lo, hi, bb in pieces if bb == b]
    assert sum(hi - lo for _, lo, hi in mine) == batch
    for s, lo, hi in mine:
      rows = np.sort(tags[tags[:, 0] == s][:, 1])
      np.testing.assert_array_equal(rows, np.arange(lo, hi))
