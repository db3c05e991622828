"""Linear (ridge) decoding model: the reference's brain_model surface on the B200 kernels.

Mirrors telluride_decoding/brain_model.py: `pearson_correlation*` (:34-91), the `BrainModel`
shim whose `evaluate` returns a dict (:129-253), `BrainModelLinearRegression` (:306-381) and
`calculate_linear_regressor_parameters_from_dataset` (:384-481).  There is no Keras here: fit
is the closed-form solve on sufficient statistics accumulated by the CUDA kernels straight from
the raw trials (the lagged design is never built), evaluate / predict run the lagged forward
kernel and the windowed Pearson kernel.
"""

import numpy as np
import torch

from telluride_decoding_b200 import brain_data
from telluride_decoding_b200 import ops


# ------------------------------------------------------------------------------------------
# Pearson correlation                                                   brain_model.py:34-91
# ------------------------------------------------------------------------------------------

def _as_tensor(v):
  if isinstance(v, torch.Tensor):
    return v
  return torch.from_numpy(np.asarray(v))


def pearson_correlation(x, y):
  """Column-wise Pearson r between two [frames, dims] arrays (brain_model.py:34-79).

  Returns a vector with one correlation per column, or all zeros when the product over columns
  of either power is <= 0 (:72-79).  Same dtype and device as the inputs (float32 in, float32
  arithmetic, like the TF ops of the reference).
  """
  x, y = _as_tensor(x), _as_tensor(y)
  assert x.shape[-1] == y.shape[-1], ('x (%s) and y (%s) do not have the same '
                                      'final dimensionality' % (tuple(x.shape), tuple(y.shape)))
  x_m = x - torch.mean(x, dim=0)
  y_m = y - torch.mean(y, dim=0)
  x_p = torch.sum(torch.square(x_m), dim=0)
  y_p = torch.sum(torch.square(y_m), dim=0)
  if bool(torch.prod(x_p) <= 0) or bool(torch.prod(y_p) <= 0):
    return 0 * x_m[0]
  return torch.sum(x_m * y_m, dim=0) / (torch.sqrt(x_p) * torch.sqrt(y_p))


def pearson_correlation_first(x, y):
  """Return the correlation of the first dimension (brain_model.py:82-85)."""
  return pearson_correlation(x, y)[0]


def pearson_correlation_second(x, y):
  """Return the correlation of the second dimension (brain_model.py:88-91)."""
  return pearson_correlation(x, y)[1]


_METRIC_COLUMNS = {'pearson_correlation_first': 0, 'pearson_correlation_second': 1}


def _metric_name_and_column(metric):
  name = metric if isinstance(metric, str) else getattr(metric, '__name__', str(metric))
  if name.startswith('cca_'):
    base = name[len('cca_'):]
  else:
    base = name
  if base not in _METRIC_COLUMNS:
    raise ValueError('Unsupported metric %s: the closed-form models report '
                     'pearson_correlation_first / _second.' % name)
  return name, _METRIC_COLUMNS[base]


def is_dataset(obj):
  return isinstance(obj, brain_data.LaggedDataset)


# ------------------------------------------------------------------------------------------
# Sufficient statistics of a dataset
# ------------------------------------------------------------------------------------------

class RegressionStats(object):
  """sum x'x, sum x'[y|1], sum [y|1]'[y|1] over the kept rows of a dataset.

  What the reference accumulates minibatch by minibatch (brain_model.py:429-444), computed in
  one pass per weighted prefix: the D x D Gram on tensor cores (float32 out), the thin products
  in float64.  Additive: `a - b` gives the statistics with b's rows removed (leave-one-out).
  """

  def __init__(self, gram, xtv, vtv):
    self.gram = gram      # [D, D] float32 (or float64 after arithmetic)
    self.xtv = xtv        # [D, Q] float64, V = [y, 1]
    self.vtv = vtv        # [Q, Q] float64

  @property
  def n(self):
    return float(self.vtv[-1, -1].item())

  @classmethod
  def from_dataset(cls, dataset, gram_options=None, workspace=None):
    passes = dataset.passes()
    groups = dataset.segment_groups() if len(passes) == 1 else None
    if groups is not None and len(groups) > 1:
      # The trials are still arriving from the host (TrialBrainData.set_trials with pinned arrays):
      # the statistics are sums over segments, so every group of segments is processed as soon as
      # the copy that completes it has finished.
      weight, keep = passes[0]
      total = None
      for event, segs in groups:
        torch.cuda.current_stream().wait_event(event)
        part = cls._one_pass(dataset.stream('x', segs, wait=False), dataset.stream('y', segs, wait=False),
                             [keep[i] for i in segs], gram_options, None)
        if total is None:
          total = part
        else:
          total.gram += part.gram
          total.xtv += part.xtv
          total.vtv += part.vtv
      dataset.ready = None
      if weight != 1.0:
        total = cls(total.gram * weight, total.xtv * weight, total.vtv * weight)
      return total
    x = dataset.stream('x')
    y = dataset.stream('y')
    total = None
    for weight, keep in passes:
      part = cls._one_pass(x, y, keep, gram_options, workspace)
      if weight != 1.0:
        part = cls(part.gram * weight, part.xtv * weight, part.vtv * weight)
      total = part if total is None else cls(total.gram + part.gram, total.xtv + part.xtv,
                                             total.vtv + part.vtv)
    if total is None:
      raise ValueError('Dataset has no complete minibatch: %r' % dataset)
    return total

  @classmethod
  def _one_pass(cls, x, y, keep, gram_options, workspace):
    if y.channels + 1 <= 16:
      g = ops.lagged_gram([x], keep, options=gram_options, workspace=workspace)
      t, v = ops.lagged_thin(x, y, keep)
    else:
      # Wide target: x'y comes out of the two-stream Gram, the ones column from thin passes.
      gz = ops.lagged_gram([x, y], keep, options=gram_options, workspace=workspace)
      d = x.width
      g = gz[:d, :d]
      tx, n1 = ops.lagged_thin(x, None, keep)
      ty, _ = ops.lagged_thin(y, None, keep)
      t = torch.cat([gz[:d, d:].to(torch.float64), tx], dim=1)
      v = torch.cat([torch.cat([gz[d:, d:].to(torch.float64), ty], dim=1),
                     torch.cat([ty.t(), n1], dim=1)], dim=0)
    return cls(g, t, v)

  def __sub__(self, other):
    return RegressionStats(self.gram.to(torch.float64) - other.gram.to(torch.float64),
                           self.xtv - other.xtv, self.vtv - other.vtv)

  def __add__(self, other):
    return RegressionStats(self.gram.to(torch.float64) + other.gram.to(torch.float64),
                           self.xtv + other.xtv, self.vtv + other.vtv)

  def augmented(self, use_offset=True):
    """(G_aug [Da, Da], C_aug [Da, Dy]) float64 with the ones column last (brain_model.py:436)."""
    d = self.gram.shape[0]
    dy = self.xtv.shape[1] - 1
    if not use_offset:
      return self.gram.to(torch.float64), self.xtv[:, :dy].contiguous()
    g = torch.empty((d + 1, d + 1), dtype=torch.float64, device=self.gram.device)
    g[:d, :d] = self.gram
    g[:d, d] = self.xtv[:, dy]
    g[d, :d] = self.xtv[:, dy]
    g[d, d] = self.vtv[dy, dy]
    c = torch.cat([self.xtv[:, :dy], self.vtv[dy:dy + 1, :dy]], dim=0)
    return g, c.contiguous()


def solve_regression(stats, lambdas, use_offset=True, method=2):
  """[(W [D, Dy], b [Dy])] float32 CUDA tensors, one per lambda (brain_model.py:447-479).

  method 2 (default): the whole sweep as one batched blocked Cholesky; 0: cuSOLVER Cholesky per
  lambda; 1: one eigendecomposition reused for every lambda."""
  g_aug, c_aug = stats.augmented(use_offset)
  sol = ops.ridge_solve(g_aug, c_aug, stats.n, lambdas, method=method)
  out = []
  for i in range(sol.shape[0]):
    if use_offset:
      out.append((sol[i, :-1, :], sol[i, -1, :]))
    else:
      out.append((sol[i], torch.zeros((sol.shape[2],), dtype=torch.float32, device=sol.device)))
  return out


def _ledoit_wolf_shrinkage(dataset, use_offset, g_aug, mean_x, mu, n):
  """Shrinkage of brain_model.py:457-465 (float64).

  The reference accumulates `sum_x2tx2` (:440-443): every minibatch is centred by the mean of all
  rows seen up to and including it, squared elementwise, and its Gram is added up -- and then uses
  only `np.sum` of that matrix, which is sum_t (sum_i z[t, i])^2 with z the squared centred design.
  Two kernels over the implicit lagged design do the same without building a row: the column sums of
  every minibatch (-> the running means), then the row-wise statistic.  The ones column of the
  augmented design centres to exactly zero and contributes nothing.  The minibatches are those of
  the un-shuffled stream; repeated epochs fall back to walking the materialised minibatches."""
  if dataset.repeat_count != 1:
    return _ledoit_wolf_shrinkage_from_batches(dataset, use_offset, g_aug, mean_x, mu, n)
  pieces = dataset.batch_pieces()
  x = dataset.stream('x')
  px = ops.Stream(x.data, x.pre, x.post, [x.seg_start[s] for s, _, _, _ in pieces],
                  [x.seg_len[s] for s, _, _, _ in pieces])
  first = [lo for _, lo, _, _ in pieces]
  keep = [hi for _, _, hi, _ in pieces]
  slot = [b for _, _, _, b in pieces]
  n_batches = dataset.total_rows // dataset.final_batch_size
  sums, rows = ops.lagged_column_sums(px, first, keep, slot, n_batches)
  running_mean = torch.cumsum(sums, dim=0) / torch.cumsum(rows, dim=0).unsqueeze(1)
  stat = ops.lagged_lw_stat(px, first, keep, slot, running_mean)
  d = g_aug.shape[0]
  cov_x_zc = g_aug - torch.outer(mean_x, mean_x)
  delta_ = cov_x_zc.clone()
  delta_.diagonal().sub_(mu)
  delta = float((delta_ ** 2).sum().item()) / d
  beta_ = (stat / n - float((cov_x_zc ** 2).sum().item())) / (d * n)
  return min(beta_, delta) / delta


def _ledoit_wolf_shrinkage_from_batches(dataset, use_offset, g_aug, mean_x, mu, n):
  """The same statistic from the materialised minibatches, in the reference's order (repeated epochs)."""
  sum_x = None
  stat = 0.0
  rows_seen = 0
  for inputs, _ in dataset:
    x = inputs['input_1'].to(torch.float64)
    if use_offset:
      x = torch.cat((x, torch.ones((x.shape[0], 1), dtype=x.dtype, device=x.device)), dim=1)
    rows_seen += x.shape[0]
    batch_sum = x.sum(dim=0, keepdim=True)
    sum_x = batch_sum if sum_x is None else sum_x + batch_sum
    centred_sq = (x - sum_x / rows_seen) ** 2
    stat += float((centred_sq.sum(dim=1) ** 2).sum().item())
  d = g_aug.shape[0]
  cov_x_zc = g_aug - torch.outer(mean_x, mean_x)
  delta_ = cov_x_zc.clone()
  delta_.diagonal().sub_(mu)
  delta = float((delta_ ** 2).sum().item()) / d
  beta_ = (stat / n - float((cov_x_zc ** 2).sum().item())) / (d * n)
  return min(beta_, delta) / delta


def calculate_linear_regressor_parameters_from_dataset(dataset, lamb=0.1, use_offset=True,
                                                       use_ridge=True, return_covariances=True):
  """Estimate the parameters for a linear regressor from a dataset (brain_model.py:384-481).

  Returns (W [D, Dy], b [1, Dy] (zeros((1,)) without offset), cov_x, cov_xy, shrinkage) as numpy
  float32 arrays like the reference.  `use_ridge=False` is the Blankertz shrinkage of :456-476
  with a fixed lambda in [0, 1] or, for lamb == -1, the Ledoit-Wolf estimate of :440-443, :457-465.
  Ledoit-Wolf needs the squared design centred by the RUNNING mean of the minibatches seen so far;
  only the sum of all entries of its Gram is used, and that comes from two kernels over the implicit
  lagged design (td_lagged_column_sums_f64, td_lagged_lw_stat_f64).
  """
  if not is_dataset(dataset):
    raise TypeError('dataset input to '
                    'calculate_linear_regressor_parameters_from_database must '
                    'be a tf.data.Dataset object')
  stats = RegressionStats.from_dataset(dataset)
  n = stats.n
  g_aug, c_aug = stats.augmented(use_offset)
  n_col = g_aug.shape[0]
  if use_ridge:
    shrinkage = lamb
    sol = ops.ridge_solve(g_aug, c_aug, n, [lamb])[0]
    reg_diag, scale = float(lamb), 1.0
  else:
    # brain_model.py:450-452: the un-normalised sum minus ONE mean outer product (sic).
    mean_x = g_aug[:, -1] / n if use_offset else stats.xtv[:, -1] / n
    trace_zc = torch.trace(g_aug) - torch.sum(mean_x * mean_x)
    mu = float(trace_zc.item()) / n_col
    if lamb == -1:
      shrinkage = _ledoit_wolf_shrinkage(dataset, use_offset, g_aug, mean_x, mu, n)
    elif lamb > 1 or lamb < 0:
      raise ValueError('Regularization lambda must be between 0 and 1, not %g.' % lamb)
    else:
      shrinkage = lamb
    if shrinkage >= 1.0:
      sol = (c_aug / n / mu).to(torch.float32)
    else:   # ((1-s) A0 + s mu I) w = B0   <=>   (A0 + s mu/(1-s) I) w = B0 / (1-s)
      sol = ops.ridge_solve(g_aug, c_aug, n, [shrinkage * mu / (1.0 - shrinkage)])[0]
      sol = sol / (1.0 - shrinkage)
    reg_diag, scale = shrinkage * mu, 1.0 - shrinkage
  cov_x = cov_xy = None
  if return_covariances:
    cov_x = (g_aug / n * scale)
    cov_x.diagonal().add_(reg_diag)
    cov_x = cov_x.to(torch.float32).cpu().numpy()
    cov_xy = (c_aug / n).to(torch.float32).cpu().numpy()
  sol = sol.cpu().numpy()
  if use_offset:
    return sol[0:-1, :], sol[-1:, :], cov_x, cov_xy, shrinkage
  return sol, np.zeros((1,)), cov_x, cov_xy, shrinkage


# ------------------------------------------------------------------------------------------
# Models
# ------------------------------------------------------------------------------------------

class BrainModel(object):
  """Shim with the reference's fit / evaluate / predict surface (brain_model.py:129-301)."""

  def __init__(self, tensorboard_dir=None, **kwargs):
    del kwargs
    self._tensorboard_dir = tensorboard_dir
    self._loss = None
    self._metrics = []
    self.metrics_names = ['loss']
    self.telluride_metadata = None

  @property
  def tensorboard_dir(self):
    return self._tensorboard_dir

  def compile(self, optimizer=None, loss=None, metrics=None, learning_rate=1e-3, **kwargs):
    """Records loss / metrics; no optimiser is ever used on the closed-form path."""
    del optimizer, learning_rate, kwargs
    self._loss = loss
    if metrics is None:
      metrics = []
    if not isinstance(metrics, (list, tuple)):
      metrics = [metrics,]
    self._metrics = [_metric_name_and_column(m) for m in metrics]
    self.metrics_names = ['loss'] + [n for n, _ in self._metrics]

  def add_metadata(self, flags, dataset=None):
    """Keeps experiment parameters with the model (brain_model.py:255-280)."""
    self.telluride_metadata = dict(flags) if isinstance(flags, dict) else flags
    if not dataset:
      return
    if not is_dataset(dataset):
      raise TypeError('dataset parameter must be tf.data.Dataset type.')
    spec_in, spec_out = dataset.element_spec
    self.telluride_inputs = {k: [dataset.final_batch_size, v.shape[-1]] for k, v in spec_in.items()}
    self.telluride_output = [dataset.final_batch_size, spec_out.shape[-1]]

  # Subclasses provide: _forward_dataset(dataset) -> [total_rows, width] (file-major rows),
  # call(input_dict) -> [batch, width].
  def _batch_metrics(self, dataset):
    raise NotImplementedError

  def evaluate(self, x=None, y=None, epoch_count=1, **kwargs):
    """Evaluate on a dataset; returns a dict of results (brain_model.py:206-253)."""
    del y, epoch_count, kwargs
    if not is_dataset(x):
      raise TypeError('BrainModel.evaluate must be called with tf.data.Dataset'
                      ' object.')
    return self._batch_metrics(x)

  def __call__(self, input_dict):
    return self.call(input_dict)


def _stream_order(dataset, rows):
  """File-major [total_rows, w] -> the kept prefix of the interleaved stream."""
  kept = dataset.num_batches * dataset.final_batch_size
  if dataset.repeat_count != 1:
    raise ValueError('evaluate expects repeat_count == 1 (test / validate datasets).')
  if len(dataset.files) == 1:
    return rows[:kept]
  return rows[dataset.stream_permutation(kept)]


def batch_pearson(dataset, truth, pred):
  """float64 [num_batches, dims]: Pearson r of every minibatch (Keras metric semantics:
  brain_model.py:62-85 evaluated per batch of `final_batch_size` stream rows)."""
  b = dataset.final_batch_size
  n = dataset.num_batches * b
  vals, _ = ops.windowed_corr(truth, pred, [0], [n], [b], [b], 'local')
  return vals


class BrainModelLinearRegression(BrainModel):
  """Linear regression computed deterministically from the data (brain_model.py:306-381)."""

  def __init__(self, input_dataset, regularization_lambda=0.0, tensorboard_dir=None,
               gram_options=None, **kwargs):
    """gram_options (not in the reference): ops.gram_options(...) for the statistics pass of fit(),
    e.g. method=5 for the block-Toeplitz path; None = the dense tcgen05 kernel."""
    super(BrainModelLinearRegression, self).__init__(tensorboard_dir=tensorboard_dir, **kwargs)
    self._gram_options = gram_options
    if not is_dataset(input_dataset):
      raise ValueError('Dataset must be a tf.data.datasert, not a %s' % type(input_dataset))
    self._input_width = input_dataset.element_spec[0]['input_1'].shape[-1]
    self._output_width = input_dataset.element_spec[1].shape[-1]
    self._regularization_lambda = regularization_lambda
    self._w = None     # [D, Dy] float32 CUDA
    self._b = None     # [Dy]
    self.compile()

  def compile(self, optimizer=None, loss='mse', metrics=pearson_correlation_first,
              learning_rate=1e-3, **kwargs):
    super(BrainModelLinearRegression, self).compile(
        optimizer=optimizer, loss=loss, metrics=metrics, learning_rate=learning_rate, **kwargs)

  def set_weights(self, weights):
    w, b = weights
    dev = self._w.device if self._w is not None else 'cuda'
    self._w = torch.as_tensor(w, dtype=torch.float32, device=dev).contiguous()
    self._b = torch.as_tensor(b, dtype=torch.float32, device=dev).reshape(-1).contiguous()
    self.w_estimate = self._w.cpu().numpy()
    self.b_estimate = self._b.cpu().numpy()

  def fit_from_stats(self, stats, regularization_lambda=None, method=2):
    lamb = self._regularization_lambda if regularization_lambda is None else regularization_lambda
    (w, b), = solve_regression(stats, [lamb], method=method)
    self._w, self._b = w.contiguous(), b.contiguous()
    self.w_estimate = self._w.cpu().numpy()
    self.b_estimate = self._b.cpu().numpy()

  def fit(self, input_dataset, y=None, **kwargs):
    """Closed-form training (brain_model.py:361-377); returns {} (no history)."""
    del kwargs
    if not is_dataset(input_dataset):
      raise TypeError('BrainModelLinearRegression.train must be called with '
                      'tf.data.Dataset, not %s.' % type(input_dataset))
    if y is not None:
      raise ValueError('Y value not needed, should be part of dataset.')
    self.fit_from_stats(RegressionStats.from_dataset(input_dataset, gram_options=self._gram_options))
    return {}

  @property
  def weight_matrices(self):
    return [self.w_estimate, self.b_estimate]

  def call(self, input_dataset):
    """Dense forward on an already-lagged batch dict (brain_model.py:340-341)."""
    x = _as_tensor(input_dataset['input_1']).to(self._w.device, torch.float32)
    return x @ self._w + self._b

  def _forward_dataset(self, dataset):
    if self._w is None:
      raise ValueError('Model has not been fit yet.')
    return ops.lagged_predict(dataset.stream('x'), dataset.seg_rows, dataset.seg_out_rows(),
                              self._w, bias=self._b, total_rows=dataset.total_rows)

  def predict(self, x, **kwargs):
    """Predictions for a batch dict (numpy in, numpy out) or for a whole dataset (stream order)."""
    del kwargs
    if is_dataset(x):
      return _stream_order(x, self._forward_dataset(x)).cpu().numpy()
    return self.call(x).cpu().numpy()

  def _batch_metrics(self, dataset):
    if dataset.num_batches == 0:
      raise ValueError('Dataset has no complete minibatch: %r' % dataset)
    pred = _stream_order(dataset, self._forward_dataset(dataset))
    truth = _stream_order(dataset, dataset.unlagged_rows('y')).contiguous()
    r = batch_pearson(dataset, truth, pred)
    results = {'loss': float(torch.mean(torch.square(truth - pred)).item())}
    for name, col in self._metrics:
      results[name] = float(torch.mean(r[:, col]).item())
    return results
