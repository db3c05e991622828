"""CPU oracle for the dense decoding hot path.  TEST INFRASTRUCTURE ONLY.

A numpy restatement of what google/telluride_decoding computes on the path
SURVEY.md section 8 scopes: lag (pre/post-context) design, interleave + batching
row set, streamed Gram / cross-covariance, ridge and CCA solves, per-batch Pearson
evaluation, the Decoder's streaming correlator, window indexing and the
winner-take-all decision.  Every function cites the reference file:line it follows
(paths relative to /root/reference/).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
legs may import this module; the product package must never do so.

Parity status: PINNED.  `oracle/gen_golden.py` runs the reference's own, unmodified
numpy functions (brain_model.py, cca.py, infer_decoder.py, result_store.py,
attention_decoder.py under the stub `tensorflow` in oracle/tfstub) on seeded inputs
and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this
restatement against those files and against the known-answer vectors in the
reference's unit tests.  The TF-only pieces (tf.signal.frame lag builder, tf.data
interleave/batch, Keras evaluate averaging) cannot execute here; they are restated
from the source and pinned by the reference's literal goldens
(test/brain_data_test.py:291-294, :317-320, :232-272; test/preprocess_test.py:250-264).
"""

import numpy as np


# ----------------------------------------------------------------------------
# Lag (temporal context) design            brain_data.py:425-483
# ----------------------------------------------------------------------------

def add_context(x, pre_context, post_context):
  """Rows [x[t-pre] .. x[t] .. x[t+post]] flattened lag-major, channel-minor.

  brain_data.py:446-455: zero rows are concatenated before (pre) and after (post),
  tf.signal.frame cuts windows of pre+1+post frames at step 1, reshape flattens each
  window so column = lag*C + channel.  Exactly N output rows; out-of-range is zero.
  """
  x = np.asarray(x)
  n, c = x.shape
  lags = pre_context + 1 + post_context
  padded = np.zeros((n + pre_context + post_context, c), dtype=x.dtype)
  padded[pre_context:pre_context + n] = x
  out = np.empty((n, lags * c), dtype=x.dtype)
  for k in range(lags):
    out[:, k * c:(k + 1) * c] = padded[k:k + n]
  return out


def window_data(x, x2, y, a, pre=0, post=0, pre2=0, post2=0, input_offset=0):
  """One chunk of the four streams -> lagged rows, zip-truncated.

  brain_data.py:459-483: positive offset drops leading rows of x only, negative
  drops them from x2 and y (not from the attention stream); x gets (pre, post),
  x2 gets (pre2, post2), y none; Dataset.zip stops at the shortest stream.
  """
  if input_offset > 0:
    x = x[input_offset:]
  elif input_offset < 0:
    x2 = x2[-input_offset:]
    y = y[-input_offset:]
  xc = add_context(x, pre, post)
  x2c = add_context(x2, pre2, post2)
  n = min(xc.shape[0], x2c.shape[0], y.shape[0], a.shape[0])
  return xc[:n], x2c[:n], np.asarray(y)[:n], np.asarray(a)[:n]


def file_rows(x, x2, y, a, pre=0, post=0, pre2=0, post2=0, input_offset=0,
              initial_batch_size=1000000, offset_counts=True):
  """All lagged rows of ONE file / saved array.

  brain_data.py:485-502: if any context (or offset) is requested the stream is cut
  into chunks of `initial_batch_size` frames and window_data runs per chunk, so the
  zero padding (and the offset drop) recurs at every chunk edge.  `offset_counts`
  False reproduces TFExampleData.read_data_into_dataset (brain_data.py:761-763),
  whose context test ignores input_offset.
  """
  x = np.asarray(x)
  x2 = np.asarray(x2)
  y = np.asarray(y)
  a = np.asarray(a)
  wants = pre or post or pre2 or post2 or (input_offset if offset_counts else 0)
  if not wants:
    return x, x2, y, a
  parts = []
  for s in range(0, x.shape[0], initial_batch_size):
    e = s + initial_batch_size
    parts.append(window_data(x[s:e], x2[s:e], y[s:e], a[s:e], pre, post, pre2,
                             post2, input_offset))
  return tuple(np.concatenate([p[i] for p in parts], axis=0) for i in range(4))


# ----------------------------------------------------------------------------
# Row set and batch boundaries            brain_data.py:720-731, :325-393
# ----------------------------------------------------------------------------

def interleave_order(lengths):
  """Stream order of tf.data interleave(cycle_length=n_files, block_length=1).

  brain_data.py:725-728.  Returns (file_index, row_index) arrays: round r emits row r
  of every file that still has one, in file order.
  """
  lengths = np.asarray(lengths, dtype=np.int64)
  files, rows = [], []
  for r in range(int(lengths.max()) if lengths.size else 0):
    alive = np.nonzero(lengths > r)[0]
    files.append(alive)
    rows.append(np.full(alive.shape, r, dtype=np.int64))
  if not files:
    return np.zeros((0,), np.int64), np.zeros((0,), np.int64)
  return np.concatenate(files), np.concatenate(rows)


def batches_from_files(per_file_rows, batch_size, repeat_count=1):
  """[(dict, y)] minibatches exactly as final_shuffle_and_batch yields them.

  brain_data.py:350-390 with shuffle_buffer_size == 0: optional repeat (train mode),
  batch(final_batch_size, drop_remainder=True), then the 4-tuple becomes
  ({'input_1','input_2','attended_speaker'}, y).
  """
  lengths = [p[0].shape[0] for p in per_file_rows]
  fi, ri = interleave_order(lengths)
  def gather(stream):
    width = per_file_rows[0][stream].shape[1]
    out = np.empty((fi.shape[0], width), dtype=per_file_rows[0][stream].dtype)
    for f, p in enumerate(per_file_rows):
      sel = fi == f
      out[sel] = p[stream][ri[sel]]
    return out
  x, x2, y, a = (gather(s) for s in range(4))
  if repeat_count > 1:
    x, x2, y, a = (np.concatenate([v] * repeat_count, axis=0) for v in (x, x2, y, a))
  nb = x.shape[0] // batch_size
  out = []
  for b in range(nb):
    s, e = b * batch_size, (b + 1) * batch_size
    out.append(({'input_1': x[s:e], 'input_2': x2[s:e], 'attended_speaker': a[s:e]},
                y[s:e]))
  return out


# ----------------------------------------------------------------------------
# Ridge / shrinkage regression            brain_model.py:384-481
# ----------------------------------------------------------------------------

def linear_regressor_from_batches(batches, lamb=0.1, use_offset=True, use_ridge=True):
  """Streamed normal equations, float32 like the reference.

  brain_model.py:429-444 accumulates x'x, sum x, x'y per batch (ones column appended
  when use_offset); :447-454 divides by the row count and adds lamb*I over ALL
  D+1 rows (bias included); :456-476 is the Blankertz / Ledoit-Wolf shrinkage
  branch; :477-481 solves and splits weights from the bias row.
  """
  sum_x = 0.0
  sum_xtx = 0.0
  sum_x2tx2 = 0
  sum_xty = 0
  n = 0
  for xd, y in batches:
    x = np.asarray(xd['input_1'])
    y = np.asarray(y)
    rows = x.shape[0]
    n += rows
    if use_offset:
      x = np.hstack((x, np.ones((rows, 1), dtype=x.dtype)))
    sum_xtx = sum_xtx + np.matmul(x.T, x)
    sum_x = sum_x + np.sum(x, axis=0, keepdims=True)
    sum_xty = sum_xty + np.matmul(x.T, y)
    if lamb == -1:
      centred_sq = (x - sum_x / n) ** 2
      sum_x2tx2 = sum_x2tx2 + np.matmul(centred_sq.T, centred_sq)
  cov_x = sum_xtx / n
  cov_xy = sum_xty / n
  mean_x = sum_x / n
  cov_x_zc = sum_xtx - np.matmul(mean_x.T, mean_x)
  d = cov_x.shape[0]
  mu = np.trace(cov_x_zc) / d
  if use_ridge:
    cov_x += lamb * np.identity(d)
    shrinkage = lamb
  else:
    if lamb == -1:
      cov_x2 = sum_x2tx2 / n
      delta_ = cov_x_zc.copy()
      delta_.flat[::d + 1] -= mu
      delta = (delta_ ** 2).sum() / d
      beta_ = 1. / (d * n) * np.sum(cov_x2 - (cov_x_zc ** 2))
      shrinkage = min(beta_, delta) / delta
    elif lamb > 1 or lamb < 0:
      raise ValueError('Regularization lambda must be between 0 and 1, not %g.' % lamb)
    else:
      shrinkage = lamb
    cov_x = (1 - shrinkage) * cov_x + shrinkage * mu * np.identity(d)
  solution = np.linalg.solve(cov_x, cov_xy)
  if use_offset:
    return solution[:-1, :], solution[-1:, :], cov_x, cov_xy, shrinkage
  return solution, np.zeros((1,)), cov_x, cov_xy, shrinkage


def sufficient_stats(batches, with_input2=False):
  """float64 sums behind both solvers (for tie-breaking 'who is closer to truth')."""
  g = c = sx = sy = gyy = 0.0
  n = 0
  for xd, y in batches:
    x = np.asarray(xd['input_1'], dtype=np.float64)
    t = np.asarray(xd['input_2'] if with_input2 else y, dtype=np.float64)
    g = g + x.T @ x
    c = c + x.T @ t
    gyy = gyy + t.T @ t
    sx = sx + x.sum(0)
    sy = sy + t.sum(0)
    n += x.shape[0]
  return dict(xtx=g, xty=c, yty=gyy, sum_x=sx, sum_y=sy, n=n)


# ----------------------------------------------------------------------------
# CCA                                      cca.py:272-369
# ----------------------------------------------------------------------------

def cca_from_batches(batches, dim, regularization=0.1, mini_batch_count=1000,
                     eps_eig=1e-12, dtype=None):
  """cca.py:311-332 streams x'x, y'y, x'y and column sums over input_1 / input_2;
  :337-343 normalises by (n_batches * last_batch_rows - 1), subtracts the mean outer
  product WITHOUT rescaling it and adds reg*I to the two auto-covariances;
  :345-366 whitens through a general eig, SVDs the whitened cross-covariance and
  returns the first `dim` directions.

  dtype=np.float64 is the FLOAT64 TWIN: the same steps on inputs promoted to
  float64.  The reference runs them in the float32 of its inputs, and a float32 general `eig`
  + SVD leaves ~1e-4 of rounding in the rotations; the twin says how far the reference's own
  output is from the exact answer of its algorithm.
  """
  if regularization < 0.0:
    raise ValueError('regularization lambda must be >= 0')
  cxx = cyy = cxy = 0
  sx = sy = 0
  nb = 0
  frames = 0
  rows = 0
  for xd, _ in (batches if not mini_batch_count else batches[:mini_batch_count]):
    x = np.asarray(xd['input_1'], dtype=dtype)
    y = np.asarray(xd['input_2'], dtype=dtype)
    if x.shape[1] == 0:
      raise ValueError('First input to CCA estimator must have more than 0 columns.')
    if y.shape[1] == 0:
      raise ValueError('Second input to CCA estimator must have more than 0 columns.')
    rows = x.shape[0]
    frames += rows
    cxx = cxx + np.matmul(x.T, x)
    cyy = cyy + np.matmul(y.T, y)
    cxy = cxy + np.matmul(x.T, y)
    sx = sx + np.sum(x, axis=0, keepdims=True)
    sy = sy + np.sum(y, axis=0, keepdims=True)
    nb += 1
  if not nb:
    raise ValueError('No minibatches in dataset, can\'t compute CCA model.')
  mean_x = sx / frames
  mean_y = sy / frames
  norm = nb * rows - 1
  cxx = cxx / norm - np.matmul(mean_x.T, mean_x)
  cxx += regularization * np.eye(cxx.shape[0])
  cyy = cyy / norm - np.matmul(mean_y.T, mean_y)
  cyy += regularization * np.eye(cyy.shape[0])
  cxy = cxy / norm - np.matmul(mean_x.T, mean_y)

  xv, xe = np.linalg.eig(cxx)
  yv, ye = np.linalg.eig(cyy)
  keep = np.where(xv > eps_eig)[0]
  xv, xe = xv[keep], xe[:, keep]
  keep = np.where(yv > eps_eig)[0]
  yv, ye = yv[keep], ye[:, keep]
  k11 = np.matmul(np.matmul(xe, np.diag(np.reciprocal(np.sqrt(xv)))), xe.transpose())
  k22 = np.matmul(np.matmul(ye, np.diag(np.reciprocal(np.sqrt(yv)))), ye.transpose())
  t = np.matmul(np.matmul(k11, cxy), k22)
  u, e, vt = np.linalg.svd(t, full_matrices=False)
  v = vt.transpose()
  return (np.matmul(k11, u[:, :dim]), np.matmul(k22, v[:, :dim]), mean_x, mean_y,
          e[:dim])


def cca_forward(x1, x2, mean1, mean2, rot1, rot2):
  """BrainCcaLayer.call, cca.py:150-161: [(x1-mean1) rot1 , (x2-mean2) rot2]."""
  return np.concatenate((np.matmul(x1 - mean1, rot1), np.matmul(x2 - mean2, rot2)),
                        axis=1)


# ----------------------------------------------------------------------------
# Pearson metric and Keras-style evaluate   brain_model.py:34-91, cca.py:39-78
# ----------------------------------------------------------------------------

def pearson_correlation(x, y):
  """Column-wise r over axis 0; all zeros when either power product is <= 0
  (brain_model.py:62-79)."""
  x = np.asarray(x)
  y = np.asarray(y)
  xm = x - np.mean(x, axis=0)
  ym = y - np.mean(y, axis=0)
  xp = np.sum(np.square(xm), axis=0)
  yp = np.sum(np.square(ym), axis=0)
  if np.prod(xp) <= 0 or np.prod(yp) <= 0:
    return 0 * xm[0]
  return np.sum(xm * ym, axis=0) / (np.sqrt(xp) * np.sqrt(yp))


def cca_pearson_correlation(y_pred):
  """cca.py:61-67: split the concatenated CCA output in halves and correlate them."""
  width = y_pred.shape[-1] // 2
  if 2 * width != y_pred.shape[-1]:
    raise ValueError('CCA y matrix does not have even # dims (%d)' % y_pred.shape[-1])
  return pearson_correlation(y_pred[:, :width], y_pred[:, width:])


def evaluate_linear(batches, w, b, column=0):
  """Keras evaluate of BrainModelLinearRegression (brain_model.py:206-253, :340-359):
  per batch Dense forward, loss = mean squared error, metric = pearson r of `column`;
  both reported as the unweighted mean over (equal-size) batches."""
  losses, metrics = [], []
  for xd, y in batches:
    pred = np.matmul(np.asarray(xd['input_1']), w) + np.reshape(b, (1, -1))
    losses.append(np.mean(np.square(np.asarray(y) - pred)))
    metrics.append(pearson_correlation(np.asarray(y), pred)[column])
  name = 'pearson_correlation_first' if column == 0 else 'pearson_correlation_second'
  return {'loss': float(np.mean(losses)), name: float(np.mean(metrics))}


def evaluate_cca(batches, mean1, mean2, rot1, rot2, column=0):
  """Keras evaluate of BrainModelCCA (cca.py:193-212): loss and metric are both the
  per-batch cca_pearson_correlation_first, averaged over batches."""
  metrics = []
  for xd, _ in batches:
    out = cca_forward(np.asarray(xd['input_1']), np.asarray(xd['input_2']), mean1,
                      mean2, rot1, rot2)
    metrics.append(cca_pearson_correlation(out)[column])
  m = float(np.mean(metrics))
  return {'loss': m, 'cca_pearson_correlation_first': m}


# ----------------------------------------------------------------------------
# Decoder correlator, reductions           infer_decoder.py:288-328, :416-455
# ----------------------------------------------------------------------------

class Correlator(object):
  """Streaming global statistics; float32 in, float32 state."""

  def __init__(self):
    self.count = 0
    self.sum_x = self.sum_y = self.sum_x2 = self.sum_y2 = 0.0
    self.mean_x = self.mean_y = 0.0
    self.power = 1.0

  def add(self, x, y):
    """infer_decoder.py:299-310."""
    self.count += x.shape[0]
    self.sum_x = self.sum_x + np.sum(x, axis=0)
    self.sum_y = self.sum_y + np.sum(y, axis=0)
    self.sum_x2 = self.sum_x2 + np.sum(x ** 2, axis=0)
    self.sum_y2 = self.sum_y2 + np.sum(y ** 2, axis=0)
    self.mean_x = self.sum_x / self.count
    self.mean_y = self.sum_y / self.count
    self.power = (np.sqrt((self.sum_x2 - self.sum_x ** 2 / self.count) *
                          (self.sum_y2 - self.sum_y ** 2 / self.count)) / self.count)

  def correlate(self, x, y):
    """infer_decoder.py:327-328."""
    return (x - self.mean_x) * (y - self.mean_y) / self.power


def reduce_correlations(c, reduction):
  """infer_decoder.py:441-455 (all but 'lda')."""
  if reduction == 'first':
    return c[:, 0]
  if reduction == 'second':
    return c[:, 1]
  if reduction == 'mean':
    return np.mean(c, axis=1)
  if reduction == 'mean-squared':
    return np.mean(np.sign(c) * c ** 2, axis=1)
  if reduction == 'all':
    return c
  raise ValueError('Unknown reduction technique: %s.' % reduction)


def average_data(data, window_size):
  """infer_decoder.py:776-783: non-overlapping window means, tail dropped."""
  if window_size <= 1:
    return data
  k = data.shape[0] // window_size
  return data[:k * window_size].reshape(k, window_size, data.shape[1]).mean(axis=1)


def calculate_dprime(d1, d2):
  """infer_decoder.py:740-745 (population variances)."""
  return (np.mean(d2) - np.mean(d1)) / np.sqrt((np.var(d1) + np.var(d2)) / 2.0)


# ----------------------------------------------------------------------------
# Window indexing and the decision        result_store.py:253-338, infer.py:247-266,
#                                          attention_decoder.py:128-173
# ----------------------------------------------------------------------------

def window_starts(n_frames, width, step):
  """Windows [i*step, i*step+width) while at least `width` frames remain
  (result_store.py:262-271) -> (n - width)//step + 1 of them."""
  if n_frames < width:
    return np.zeros((0,), np.int64)
  return np.arange((n_frames - width) // step + 1, dtype=np.int64) * step


def windowed_means(signal, label, width):
  """regress_and_correlate (infer.py:260-266) over test_by_window
  (infer_decoder.py:498-504): step = width//2, float64 store, np.mean per window."""
  signal = np.asarray(signal, dtype=np.float64).reshape(-1)
  label = np.asarray(label, dtype=np.float64).reshape(-1)
  starts = window_starts(signal.shape[0], width, width // 2)
  res = np.array([np.mean(signal[s:s + width]) for s in starts])
  lab = np.array([np.mean(label[s:s + width]) for s in starts])
  return res, lab


def wta_attention(r1, r2):
  """attention_decoder.py:134: strict greater-than of the window means."""
  return np.asarray(r1) > np.asarray(r2)


def step_attention(r1, r2):
  """attention_decoder.py:153-173: +-0.1 random walk clamped to [0.1, 0.9], start 0.5."""
  state = 0.5
  out = []
  for a, b in zip(r1, r2):
    if a > b:
      state = min(0.9, state + 0.1)
    else:
      state = max(0.1, state - 0.1)
    out.append(state > 0.5)
  return np.array(out, dtype=bool)


def aad_accuracy(attention, labels):
  """infer.py:399-407 verbatim semantics: correct = xor(attention >= 0.5, label)."""
  att = np.asarray(attention, dtype=np.float64).reshape(-1, 1)
  lab = np.reshape(np.asarray(labels), (-1, 1))
  correct = np.logical_xor(att >= 0.5, lab)
  return np.sum(correct) / float(len(correct))


# ----------------------------------------------------------------------------
# Streaming context                        preprocess.py:487-527
# ----------------------------------------------------------------------------

class StreamingContext(object):
  """Online lag builder: Preprocessor.add_context (preprocess.py:487-522).  The state is the last
  pre+post frames (zeros of length pre before the first batch, :500-501); each call returns the
  rows of the running stream that are complete, i.e. delayed by post_context frames."""

  def __init__(self, pre_context, post_context):
    self.pre, self.post = int(pre_context), int(post_context)
    self.state = None

  def reset(self):
    self.state = None

  def add_context(self, data):
    data = np.asarray(data)
    if self.pre == 0 and self.post == 0:
      return data
    if self.state is None:
      self.state = np.zeros((self.pre, data.shape[1]), dtype=data.dtype)
    joined = np.concatenate((self.state, data))
    self.state = joined[-(self.pre + self.post):, :]
    rows = joined.shape[0] - self.pre - self.post
    if rows < 0:
      raise ValueError('negative dimensions are not allowed')
    lags = self.pre + self.post + 1
    out = np.empty((rows, data.shape[1] * lags), dtype=data.dtype)
    for k in range(lags):                                     # block k = joined[t + k]
      out[:, k * data.shape[1]:(k + 1) * data.shape[1]] = joined[k:k + rows]
    return out
