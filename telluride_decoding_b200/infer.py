"""Attention-decoding test over window sizes (telluride_decoding/infer.py:327-464).

`run_reduction_test` takes the decoder object and the two speakers' test datasets instead of
directories of SavedModels / TFRecords, evaluates every window size in one pass, applies the
attention decoder and reports the fraction of correct windows exactly as infer.py:399-407.
"""

import re

import numpy as np
import torch

from telluride_decoding_b200 import attention_decoder
from telluride_decoding_b200 import infer_decoder
from telluride_decoding_b200 import ops

DEFAULT_WINDOW_LIST = [10, 100, 200, 400, 700, 1000]      # infer.py:376


def find_first_segment(labels):
  """Frame at the end of the first run of equal labels (infer.py:301-324)."""
  if isinstance(labels, list):
    labels = np.asarray(labels)
  if not isinstance(labels, np.ndarray):
    raise TypeError('Labels input must be an ndarray, not %s' % type(labels))
  if labels.ndim != 1:
    raise TypeError('Labels input must be one-dimensional, not %s' % str(labels.shape))
  end_section = np.nonzero(np.logical_xor(labels, labels[0]))
  if end_section[0].shape[0]:
    return end_section[0][0]
  return 0


def calculate_time_axis(data, window_step, window_width, frame_rate):
  """Centre time (minutes) of each window (infer.py:172-199)."""
  if isinstance(data, (int, float)):
    num_points = int(data)
  elif isinstance(data, (list, np.ndarray)):
    num_points = len(data)
  else:
    raise TypeError('Unknown type passed as input argument.')
  return (np.arange(num_points) * window_step + window_width / 2.0) / frame_rate / 60.0


def run_reduction_test(model_object, bd1_test, bd2_test, decoder_type='wta',
                       window_list=None, frame_rate=100.0):
  """{window size: fraction of correctly decoded windows} (infer.py:327-464).

  bd1_test / bd2_test: test datasets whose output stream is speaker 1 / speaker 2's audio
  feature; the attention labels ride in the attended_speaker stream of bd2_test."""
  window_list = list(window_list or DEFAULT_WINDOW_LIST)
  res1 = model_object.windowed_results(bd1_test, window_list)
  res2 = model_object.windowed_results(bd2_test, window_list)
  out = {}
  for w in window_list:
    d1, _ = res1[w]
    d2, labels = res2[w]
    if decoder_type == 'wta' and len(d1):
      import torch
      att = ops.wta_decide(torch.as_tensor(d1, device='cuda'),
                           torch.as_tensor(d2, device='cuda')).cpu().numpy().astype(bool)
    else:
      decoder = attention_decoder.create_attention_decoder(decoder_type, window_step=w // 2,
                                                           frame_rate=frame_rate)
      end_first = find_first_segment(np.asarray(labels)) if len(labels) else 0
      if end_first:
        decoder.tune(d1[:end_first], d2[:end_first])
      att = np.array([decoder.attention(c1, c2)[0] for c1, c2 in zip(d1, d2)], dtype=bool)
    correct = np.logical_xor(att.reshape(-1, 1) >= 0.5, np.reshape(labels, (-1, 1)))
    out[w] = float(np.sum(correct)) / float(max(len(correct), 1))
  return out


def leave_one_trial_out_test(brain_data_obj, regularization_lambda=0.1, window_list=None,
                             reduction='first', decoder_type='wta', frame_rate=100.0,
                             fold_group=64, return_details=False, gram_options=None):
  """Two-speaker attention decoding with a leave-one-trial-out linear decoder
  (BASELINE.json configs[4]): {window size: fraction of correctly decoded windows}.

  `brain_data_obj` (TrialBrainData / TFExampleData, one file per trial):
      input_1 = EEG / MEG with the decoder's temporal context, output = the ATTENDED speaker's audio
      feature (the training target), input_2 = speaker 1's and speaker 2's features side by side
      (2 * Dy columns, no context), attended_speaker = the label (0: speaker 1 attended).

  Per held-out trial f the reference would run one complete fit on the other trials
  (decoding.train_and_test via regression.jackknife_one_model, regression.py:197-214) and then
  infer.run_reduction_test (:327-464) on f.  Here the data is swept ONCE:
    1. per-trial sufficient statistics from one launch of the tcgen05 Gram kernel (+ one of the
       thin-product kernel); fold f = total - trial f (- the rows drop_remainder removes from that
       fold's interleaved stream, brain_data.py:369-370);
    2. all fold systems from ONE Cholesky of the total system + a few fixed-point steps over the
       per-trial Grams (a held-out trial is a small perturbation of the total); the batched blocked
       Cholesky (groups of `fold_group` folds) when rows are dropped or a fold does not converge;
    3. every trial predicted by the model that never saw it (per-segment weights);
    4. correlator statistics (infer_decoder.py:288-310) pooled over both speakers' (truth,
       prediction) pairs of those held-out predictions; per-frame correlation signal, reduction,
       windows of every size inside each trial (result_store.py:262-271, step W // 2), decision
       (attention_decoder.py:128-173) and accuracy (infer.py:399-407).
  gram_options: ops.gram_options(...) for step 1 (method=5: the block-Toeplitz structure per trial).
  """
  window_list = [int(w) for w in (window_list or DEFAULT_WINDOW_LIST)]
  saved = (brain_data_obj.train_file_pattern, brain_data_obj.validate_file_pattern,
           brain_data_obj.test_file_pattern)
  brain_data_obj.set_file_patterns('', '', '')
  try:
    ds = brain_data_obj.create_dataset('test')
  finally:
    brain_data_obj.set_file_patterns(*saved)
  if ds.pre2 or ds.post2:
    raise ValueError('input_2 (the two speakers\' features) must not have temporal context')
  n_files = len(ds.files)
  if n_files < 2:
    raise ValueError('leave-one-trial-out needs at least two trials')
  x, y = ds.stream('x'), ds.stream('y')
  d, dy = x.width, y.channels
  if ds.x2.shape[1] != 2 * dy:
    raise ValueError('input_2 must hold speaker 1 and speaker 2 side by side (%d columns, not %d)' %
                     (2 * dy, ds.x2.shape[1]))
  rows = list(ds.seg_rows)
  seg_file = list(ds.seg_file)
  device = x.data.device

  # ---- 1. per-trial statistics from one pass ----------------------------------------------------
  gram_s = ops.lagged_gram_per_segment([x], rows, options=gram_options)   # [n_seg, D, D] float32
  xtv_s, vtv_s = ops.lagged_thin_per_segment(x, y, rows)             # float64
  if n_files != len(rows):                                           # several chunks per file: add them up
    idx = torch.as_tensor(seg_file, device=device)
    def per_file(t):
      out = torch.zeros((n_files,) + tuple(t.shape[1:]), dtype=torch.float64, device=device)
      return out.index_add_(0, idx, t.to(torch.float64))
    gram_f, xtv_f, vtv_f = per_file(gram_s), per_file(xtv_s), per_file(vtv_s)
  else:
    gram_f, xtv_f, vtv_f = gram_s, xtv_s, vtv_s
  del gram_s
  total_g = ops.sum_batches(gram_f)
  total_t, total_v = xtv_f.sum(dim=0), vtv_f.sum(dim=0)
  file_rows = list(ds.file_rows)
  all_rows = sum(file_rows)
  b = ds.final_batch_size

  # ---- 2. fold systems ------------------------------------------------------------------------------
  weights = torch.empty((n_files, d, dy), dtype=torch.float32, device=device)
  biases = torch.empty((n_files, dy), dtype=torch.float32, device=device)
  names = [f[0] for f in ds.files]
  solver = 'batched cholesky'
  drops = [(all_rows - file_rows[f]) % b != 0 for f in range(n_files)]
  done = False
  if not any(drops) and dy <= 4 and isinstance(gram_f, torch.Tensor) and gram_f.dtype == torch.float32:
    # ONE factorisation of the total system, then a few fixed-point steps over the per-trial Grams:
    # (S - E_f) w = c_f with E_f (one trial of n_files) small against S (ops.ridge_loo_solve).
    g_tot = torch.empty((d + 1, d + 1), dtype=torch.float64, device=device)
    g_tot[:d, :d] = total_g
    g_tot[:d, d] = total_t[:, dy]
    g_tot[d, :d] = total_t[:, dy]
    g_tot[d, d] = total_v[dy, dy]
    c_tot = torch.cat([total_t[:, :dy], total_v[dy:dy + 1, :dy]], dim=0)
    # lambda multiplies n_f * I in the un-normalised fold system (brain_model.py:447-452 divides by n_f)
    sol, iters, worst = ops.ridge_loo_solve(g_tot, c_tot, gram_f, xtv_f, vtv_f, regularization_lambda)
    if worst < 1e-8:
      weights.copy_(sol[:, :d, :])
      biases.copy_(sol[:, d, :])
      solver = 'one factorisation + %d fixed-point steps' % iters
      done = True
    del g_tot, sol
  for g0 in range(0, 0 if done else n_files, fold_group):
    g1 = min(n_files, g0 + fold_group)
    ng = g1 - g0
    g_aug = torch.empty((ng, d + 1, d + 1), dtype=torch.float64, device=device)
    g_aug[:, :d, :d] = total_g.unsqueeze(0) - gram_f[g0:g1].to(torch.float64)
    ones_col = total_t[:, dy].unsqueeze(0) - xtv_f[g0:g1, :, dy]
    g_aug[:, :d, d] = ones_col
    g_aug[:, d, :d] = ones_col
    g_aug[:, d, d] = total_v[dy, dy] - vtv_f[g0:g1, dy, dy]
    c_aug = torch.empty((ng, d + 1, dy), dtype=torch.float64, device=device)
    c_aug[:, :d, :] = total_t[:, :dy].unsqueeze(0) - xtv_f[g0:g1, :, :dy]
    c_aug[:, d, :] = total_v[dy, :dy].unsqueeze(0) - vtv_f[g0:g1, dy, :dy]
    for j in range(ng):
      f = g0 + j
      if drops[f]:
        # this fold's interleaved training stream loses its last rows to drop_remainder
        brain_data_obj.set_file_patterns('allbut', '^%s$' % _re_escape(names[f]), '^%s$' % _re_escape(names[f]))
        try:
          train_ds = brain_data_obj.create_dataset('train')
        finally:
          brain_data_obj.set_file_patterns(*saved)
        kept = train_ds.num_batches * b
        first = train_ds.seg_keep_for_prefix(kept)
        last = list(train_ds.seg_rows)
        xs = train_ds.stream('x')
        gd = ops.lagged_gram_range([xs], first, last).to(torch.float64)
        td_, vd = ops.lagged_thin(xs, train_ds.stream('y'), last, seg_first=first)
        g_aug[j, :d, :d] -= gd
        g_aug[j, :d, d] -= td_[:, dy]
        g_aug[j, d, :d] -= td_[:, dy]
        g_aug[j, d, d] -= vd[dy, dy]
        c_aug[j, :d, :] -= td_[:, :dy]
        c_aug[j, d, :] -= vd[dy, :dy]
    n_fold = [float(v) for v in g_aug[:, d, d].cpu().tolist()]
    sol = ops.ridge_solve_batched(g_aug, c_aug, n_fold, [regularization_lambda] * ng)
    weights[g0:g1] = sol[:, :d, :]
    biases[g0:g1] = sol[:, d, :]
    del g_aug, c_aug, sol

  # ---- 3. held-out predictions ----------------------------------------------------------------------
  seg_w = weights if n_files == len(rows) else weights[torch.as_tensor(seg_file, device=device)]
  seg_b = biases if n_files == len(rows) else biases[torch.as_tensor(seg_file, device=device)]
  pred = ops.lagged_predict_per_segment(x, rows, ds.seg_out_rows(), seg_w, seg_b,
                                        total_rows=ds.total_rows)
  speakers = ds.unlagged_rows('x2')
  truth1, truth2 = speakers[:, :dy].contiguous(), speakers[:, dy:].contiguous()
  labels = ds.unlagged_rows('a').contiguous()

  # ---- 4. correlator statistics, windows, decision ---------------------------------------------------
  dec = infer_decoder.LinearRegressionDecoder(None, reduction=reduction)
  dec.add_statistics((truth1, pred), (truth2, pred))
  mu_a, mu_b, power = dec._device_statistics(dy, device)   # pylint: disable=protected-access
  starts, cur = [], 0
  for n in file_rows:
    starts.append(cur)
    cur += n
  steps = [max(w // 2, 1) for w in window_list]
  r1, off = ops.windowed_corr(truth1, pred, starts, file_rows, window_list, steps, 'global',
                              reduction=reduction, mu_a=mu_a, mu_b=mu_b, power=power)
  r2, _ = ops.windowed_corr(truth2, pred, starts, file_rows, window_list, steps, 'global',
                            reduction=reduction, mu_a=mu_a, mu_b=mu_b, power=power)
  labs, _ = ops.windowed_corr(labels, None, starts, file_rows, window_list, steps, 'plain')
  if decoder_type == 'wta':
    att_all = ops.wta_decide(r1, r2).cpu().numpy().astype(bool)
  r1h, r2h, labh = r1.cpu().numpy()[:, 0], r2.cpu().numpy()[:, 0], labs.cpu().numpy()[:, 0]
  out, details = {}, {}
  for i, w in enumerate(window_list):
    lo, hi = off[i], off[i + 1]
    d1, d2, lab = r1h[lo:hi], r2h[lo:hi], labh[lo:hi]
    if decoder_type == 'wta':
      att = att_all[lo:hi]
    else:
      decoder = attention_decoder.create_attention_decoder(decoder_type, window_step=steps[i],
                                                           frame_rate=frame_rate)
      end_first = find_first_segment(np.asarray(lab)) if len(lab) else 0
      if end_first:
        decoder.tune(d1[:end_first], d2[:end_first])
      att = np.array([decoder.attention(c1, c2)[0] for c1, c2 in zip(d1, d2)], dtype=bool)
    correct = np.logical_xor(att.reshape(-1, 1) >= 0.5, np.reshape(lab, (-1, 1)))
    out[w] = float(np.sum(correct)) / float(max(len(correct), 1))
    details[w] = dict(attention=att, r1=d1, r2=d2, labels=lab)
  if return_details:
    details['weights'] = weights
    details['biases'] = biases
    details['prediction'] = pred
    details['correlator'] = dec.correlation_params
    details['solver'] = solver
    return out, details
  return out


def _re_escape(name):
  return re.escape(name)
