"""Canonical correlation analysis: the reference's cca surface on the B200 kernels.

Mirrors telluride_decoding/cca.py: `cca_pearson_correlation*` (:39-78), `BrainCcaLayer` forward
(:150-161), `BrainModelCCA` (:169-244) and `calculate_cca_parameters_from_dataset` (:272-369).
x'x, y'y and x'y over the two lagged inputs come from ONE pass of the two-stream tensor-core
Gram over Z = [X | Y]; the whitening / SVD runs in cuSOLVER (float64).
"""

import torch

from telluride_decoding_b200 import brain_model
from telluride_decoding_b200 import ops


def cca_pearson_correlation(x, y):
  """Correlation between the two halves of a CCA output (cca.py:39-67); x is ignored."""
  del x
  y = brain_model._as_tensor(y)   # pylint: disable=protected-access
  cca_width = y.shape[-1] // 2
  if 2 * cca_width != y.shape[-1]:
    raise ValueError('CCA y matrix does not have even # dims (%d)' % y.shape[-1])
  return brain_model.pearson_correlation(y[:, :cca_width], y[:, cca_width:])


def cca_pearson_correlation_first(x, y):
  return cca_pearson_correlation(x, y)[0]


def cca_pearson_correlation_second(x, y):
  return cca_pearson_correlation(x, y)[1]


class CcaStats(object):
  """Raw sums over the kept rows: [X|Y]'[X|Y], column sums, frame and batch counts
  (cca.py:311-332).  Additive, so leave-one-trial-out statistics are differences."""

  def __init__(self, gram, sum_x, sum_y, frames, d1):
    self.gram = gram          # [D1+D2, D1+D2] float32 / float64
    self.sum_x = sum_x        # [D1] float64
    self.sum_y = sum_y        # [D2] float64
    self.frames = frames
    self.d1 = d1

  @classmethod
  def from_dataset(cls, dataset, mini_batch_count=0, gram_options=None, workspace=None):
    x = dataset.stream('x')
    y = dataset.stream('x2')
    passes = dataset.passes()
    if mini_batch_count and mini_batch_count < dataset.num_batches:
      # cca.py:311: only the first mini_batch_count batches of the stream.
      kept = mini_batch_count * dataset.final_batch_size
      passes = []
      epochs, rest = divmod(kept, dataset.total_rows)
      if epochs:
        passes.append((float(epochs), list(dataset.seg_rows)))
      if rest:
        passes.append((1.0, dataset.seg_keep_for_prefix(rest)))
    gram = sx = sy = None
    frames = 0.0
    for weight, keep in passes:
      g = ops.lagged_gram([x, y], keep, options=gram_options, workspace=workspace)
      tx, n1 = ops.lagged_thin(x, None, keep)
      ty, _ = ops.lagged_thin(y, None, keep)
      if gram is None:
        gram = g if weight == 1.0 else g * weight
        sx, sy = tx[:, 0] * weight, ty[:, 0] * weight
      else:
        gram = gram + g * weight
        sx, sy = sx + tx[:, 0] * weight, sy + ty[:, 0] * weight
      frames += weight * float(n1.item())
    if gram is None:
      raise ValueError('No minibatches in dataset, can\'t compute CCA model.')
    return cls(gram, sx, sy, frames, x.width)

  def __sub__(self, other):
    return CcaStats(self.gram.to(torch.float64) - other.gram.to(torch.float64),
                    self.sum_x - other.sum_x, self.sum_y - other.sum_y,
                    self.frames - other.frames, self.d1)


def solve_cca(stats, dim, regularization=0.1, eps_eig=1e-12, n_norm=None):
  """(rot_x, rot_y, mean_x, mean_y, e) float32 CUDA tensors from CcaStats (cca.py:337-367).

  n_norm = n_batches * rows_of_last_batch - 1 (:339), which equals frames - 1 for a batched
  dataset."""
  d1 = stats.d1
  g = stats.gram
  if n_norm is None:
    n_norm = stats.frames - 1
  return ops.cca_solve(g[:d1, :d1], g[d1:, d1:], g[:d1, d1:], stats.sum_x, stats.sum_y,
                       stats.frames, n_norm, regularization, eps_eig, dim)


def calculate_cca_parameters_from_dataset(dataset, dim, regularization=0.1,
                                          mini_batch_count=1000, eps_eig=1e-12):
  """Estimate the CCA rotations from a dataset (cca.py:272-369).

  Returns numpy float32 (rot_x [D1, dim], rot_y [D2, dim], mean_x [1, D1], mean_y [1, D2],
  e [dim]).  Column signs are normalised jointly per component (largest |rot_x| entry positive).
  """
  if not brain_model.is_dataset(dataset):
    raise TypeError('dataset input to calculate_regressor_from_database must be'
                    ' a tf.data.Dataset object, not %s' % type(dataset))
  if regularization < 0.0:
    raise ValueError('regularization lambda must be >= 0')
  if dataset.element_spec[0]['input_1'].shape[-1] == 0:
    raise ValueError('First input to CCA estimator must have more '
                     'than 0 columns.')
  if dataset.element_spec[0]['input_2'].shape[-1] == 0:
    raise ValueError('Second input to CCA estimator must have more '
                     'than 0 columns.')
  if dataset.num_batches == 0:
    raise ValueError('No minibatches in dataset, can\'t compute CCA model.')
  stats = CcaStats.from_dataset(dataset, mini_batch_count=mini_batch_count)
  d2 = stats.gram.shape[0] - stats.d1
  k = min(dim, stats.d1, d2)     # numpy slicing u[:, 0:dim] clamps the same way
  rot_x, rot_y, mean_x, mean_y, e = solve_cca(stats, k, regularization, eps_eig)
  return tuple(t.cpu().numpy() for t in (rot_x, rot_y, mean_x, mean_y, e))


class BrainCcaLayer(object):
  """Holds the four CCA weights and applies the rotation (cca.py:84-166)."""

  def __init__(self, requested_cca_dims, **kwargs):
    del kwargs
    self.output_dims = requested_cca_dims
    self.mean1 = self.mean2 = self.rot1 = self.rot2 = None

  def set_initial_weights(self, mean1, mean2, rot1, rot2, device='cuda'):
    t = lambda v: torch.as_tensor(v, dtype=torch.float32, device=device).contiguous()
    self.mean1, self.mean2, self.rot1, self.rot2 = t(mean1), t(mean2), t(rot1), t(rot2)
    self.mean1 = self.mean1.reshape(1, -1)
    self.mean2 = self.mean2.reshape(1, -1)

  def get_weights(self):
    return [w.cpu().numpy() for w in (self.mean1, self.mean2, self.rot1, self.rot2)]

  def call(self, inputs):
    """[(x1 - mean1) rot1 , (x2 - mean2) rot2] on already-lagged batches (cca.py:150-161)."""
    x1 = brain_model._as_tensor(inputs[0]).to(self.rot1.device, torch.float32)   # pylint: disable=protected-access
    x2 = brain_model._as_tensor(inputs[1]).to(self.rot1.device, torch.float32)   # pylint: disable=protected-access
    return torch.cat(((x1 - self.mean1) @ self.rot1, (x2 - self.mean2) @ self.rot2), dim=1)

  __call__ = call

  def get_config(self):
    return {'requested_cca_dims': self.output_dims}


class BrainModelCCA(brain_model.BrainModel):
  """A BrainModel class that implements canonical correlation (cca.py:169-244)."""

  def __init__(self, input_dataset, cca_dims=5, regularization_lambda=0.0, **kwargs):
    super(BrainModelCCA, self).__init__(**kwargs)
    self._cca_dims = cca_dims
    self._regularization_lambda = regularization_lambda
    self._input1_width = input_dataset.element_spec[0]['input_1'].shape[-1]
    self._input2_width = input_dataset.element_spec[0]['input_2'].shape[-1]
    if self._input1_width <= 1:
      raise ValueError('Input 1 feature width (%d) should not be <= 1.' % self._input1_width)
    if self._input2_width <= 1:
      raise ValueError('Input 2 feature width (%d) should not be <= 1.' % self._input2_width)
    self._cca_layer = BrainCcaLayer(self._cca_dims)
    self.compile()

  def compile(self, optimizer=None, loss=cca_pearson_correlation_first,
              metrics=cca_pearson_correlation_first, learning_rate=1e-3, **kwargs):
    super(BrainModelCCA, self).compile(optimizer=optimizer, loss=loss, metrics=metrics,
                                       learning_rate=learning_rate, **kwargs)

  def call(self, input_data):
    return self._cca_layer([input_data['input_1'], input_data['input_2']])

  def fit_from_stats(self, stats, regularization_lambda=None):
    lamb = self._regularization_lambda if regularization_lambda is None else regularization_lambda
    d2 = stats.gram.shape[0] - stats.d1
    k = min(self._cca_dims, stats.d1, d2)
    rot_x, rot_y, mean_x, mean_y, e = solve_cca(stats, k, lamb)
    self._cca_layer.set_initial_weights(mean_x, mean_y, rot_x, rot_y, device=rot_x.device)
    self.rot_x, self.rot_y = rot_x.cpu().numpy(), rot_y.cpu().numpy()
    self.mean_x, self.mean_y = mean_x.cpu().numpy(), mean_y.cpu().numpy()
    self.eigenvalues = e.cpu().numpy()

  def fit(self, dataset, epochs=1, y=None):
    """Closed-form CCA training (cca.py:214-244); returns {}."""
    del epochs
    if not brain_model.is_dataset(dataset):
      raise TypeError('BrainModelLinearRegression.train must be called with '
                      'tf.data.Dataset.')
    if y is not None:
      raise ValueError('Y value not needed, should be part of dataset.')
    if dataset.num_batches == 0:
      raise ValueError('No minibatches in dataset, can\'t compute CCA model.')
    self.fit_from_stats(CcaStats.from_dataset(dataset, mini_batch_count=0))
    return {}

  @property
  def weight_matrices(self):
    return self._cca_layer.get_weights()

  def _forward_dataset(self, dataset):
    layer = self._cca_layer
    if layer.rot1 is None:
      raise ValueError('Model has not been fit yet.')
    k = layer.rot1.shape[1]
    out = torch.empty((dataset.total_rows, 2 * k), dtype=torch.float32, device=layer.rot1.device)
    rows, seg_out = dataset.seg_rows, dataset.seg_out_rows()
    ops.lagged_predict(dataset.stream('x'), rows, seg_out, layer.rot1,
                       mean=layer.mean1.reshape(-1), out=out[:, :k])
    ops.lagged_predict(dataset.stream('x2'), rows, seg_out, layer.rot2,
                       mean=layer.mean2.reshape(-1), out=out[:, k:])
    return out

  def predict(self, x, **kwargs):
    del kwargs
    if brain_model.is_dataset(x):
      return brain_model._stream_order(x, self._forward_dataset(x)).cpu().numpy()   # pylint: disable=protected-access
    return self.call(x).cpu().numpy()

  def _batch_metrics(self, dataset):
    if dataset.num_batches == 0:
      raise ValueError('Dataset has no complete minibatch: %r' % dataset)
    out = brain_model._stream_order(dataset, self._forward_dataset(dataset))   # pylint: disable=protected-access
    k = out.shape[1] // 2
    r = brain_model.batch_pearson(dataset, out[:, :k], out[:, k:])
    results = {}
    for name, col in self._metrics:
      results[name] = float(torch.mean(r[:, col]).item())
    # Keras reports the loss first; here it is the same correlation (cca.py:196-198).
    first = float(torch.mean(r[:, 0]).item())
    return dict([('loss', first)] + list(results.items()))
