"""Inference pipeline: decode, correlate with global statistics, reduce, window
(telluride_decoding/infer_decoder.py).

`Decoder` keeps the reference's streaming methods (`add_data_correlator` :288-310,
`compute_correlation` :312-328, `infer_one` :416-455, `test_all`, `test_by_window` :484-504) for
callers that feed minibatches, and adds the whole-dataset forms the B200 path uses:
`train_statistics(dataset)` (one pass: lagged forward kernel + column-moment kernel) and
`windowed_results(dataset, window_sizes)` (forward kernel, then ONE pass of the windowed
correlation kernel for every window size).  `average_data` (:748-783) and `calculate_dprime`
(:717-745) are here too; `train` / `compute_lda_model` / `reduce_with_lda` give the `lda` reducer
(scaled_lda.py).  SavedModel loading is outside the hot path.
"""

import collections

import numpy as np
import torch

from telluride_decoding_b200 import brain_model
from telluride_decoding_b200 import ops
from telluride_decoding_b200 import result_store
from telluride_decoding_b200 import scaled_lda

CorrelationParamsTuple = collections.namedtuple('CorrelationParamsTuple', [
    'count', 'sum_x', 'sum_y', 'sum_x2', 'sum_y2', 'mean_x', 'mean_y', 'power'
])

ModelParamsTuple = collections.namedtuple('ModelParamsTuple',
                                          ['correlation_params', 'lda_params'])

_REDUCTIONS = ('mean-squared', 'first', 'second', 'lda', 'all', 'mean')


def _to_numpy(v):
  if isinstance(v, torch.Tensor):
    return v.detach().cpu().numpy()
  if hasattr(v, 'numpy') and not isinstance(v, np.ndarray):
    return v.numpy()
  return np.asarray(v)


class Decoder(object):
  """Generic decoder: model outputs -> per-frame correlation signal -> reduced scalar."""

  def __init__(self, decoding_model=None, reduction='mean-squared'):
    if decoding_model is not None and not callable(decoding_model):
      raise TypeError('Must supply a callable model when initializing a '
                      'Decoder, not a %s.' % type(decoding_model))
    if reduction not in _REDUCTIONS:
      raise ValueError('Unknown reduction technique: %s' % reduction)
    self._decoding_model = decoding_model
    self._reduction = reduction
    self._lda = None
    self.reset_correlation_statistics()

  @property
  def decoding_model(self):
    return self._decoding_model

  @property
  def correlation_params(self):
    return CorrelationParamsTuple(self._count, self._sum_x, self._sum_y, self._sum_x2,
                                  self._sum_y2, self._mean_x, self._mean_y, self._power)

  def _set_correlation_params(self, values):
    values = CorrelationParamsTuple(*values)
    (self._count, self._sum_x, self._sum_y, self._sum_x2, self._sum_y2, self._mean_x, self._mean_y,
     self._power) = values

  @property
  def lda_params(self):
    if self._lda is None:
      self._lda = scaled_lda.ScaledLinearDiscriminantAnalysis()
    return self._lda.model_parameters

  def _set_lda_params(self, values):
    if self._lda is None:
      self._lda = scaled_lda.ScaledLinearDiscriminantAnalysis()
    self._lda.model_parameters = values

  @property
  def model_params(self):
    """Correlation statistics and LDA parameters (infer_decoder.py:197-219)."""
    return ModelParamsTuple(self.correlation_params, self.lda_params)

  @model_params.setter
  def model_params(self, values):
    self._set_correlation_params(values.correlation_params)
    self._set_lda_params(values.lda_params)

  def reset_correlation_statistics(self):
    self._count = 0
    self._sum_x = 0.0
    self._sum_y = 0.0
    self._sum_x2 = 0.0
    self._sum_y2 = 0.0
    self._mean_x = 0.0
    self._mean_y = 0.0
    self._power = 1.0

  # ------------------------------------------------------------- streaming (reference) interface
  def _finish_statistics(self):
    """mean and power from the sums (infer_decoder.py:306-310), float32 like the model outputs."""
    self._mean_x = self._sum_x / self._count
    self._mean_y = self._sum_y / self._count
    self._power = (np.sqrt((self._sum_x2 - self._sum_x ** 2 / self._count) *
                           (self._sum_y2 - self._sum_y ** 2 / self._count)) / self._count)

  def add_data_correlator(self, x, y):
    """Adds a batch to the running statistics (infer_decoder.py:288-310)."""
    x, y = _to_numpy(x), _to_numpy(y)
    self._count += x.shape[0]
    self._sum_x += np.sum(x, axis=0)
    self._sum_y += np.sum(y, axis=0)
    self._sum_x2 += np.sum(x ** 2, axis=0)
    self._sum_y2 += np.sum(y ** 2, axis=0)
    self._finish_statistics()

  def compute_correlation(self, x, y):
    """Normalised cross product without the final sum (infer_decoder.py:312-328)."""
    x, y = _to_numpy(x), _to_numpy(y)
    return ((x - np.broadcast_to(self._mean_x, x.shape)) *
            (y - np.broadcast_to(self._mean_y, y.shape)) / self._power)

  def decode_one(self, input_dict, ground_truth):
    raise NotImplementedError('Must be implemented by a subclass.')

  def _reduce(self, correlations):
    if self._reduction == 'first':
      return correlations[:, 0]
    if self._reduction == 'second':
      return correlations[:, 1]
    elif self._reduction == 'mean':
      return np.mean(correlations, axis=1)
    elif self._reduction == 'mean-squared':
      return np.mean(np.sign(correlations) * correlations ** 2, axis=1)
    elif self._reduction == 'lda':
      return self.reduce_with_lda(correlations)[:, 0]
    elif self._reduction == 'all':
      return correlations
    raise ValueError('Unknown reduction technique: %s.' % self._reduction)

  def infer_one(self, input_dict, output):
    """One minibatch through the whole pipeline (infer_decoder.py:416-455)."""
    result1, result2 = self.decode_one(input_dict, output)
    return self._reduce(self.compute_correlation(result1, result2))

  def test_all(self, exp_data):
    predictions = result_store.NumpyStore(name='test_all predictions')
    labels = result_store.NumpyStore(name='test_all labels')
    for input_dict, output in exp_data:
      predictions.add_data(self.infer_one(input_dict, output))
      labels.add_data(_to_numpy(input_dict['attended_speaker']))
    return predictions.all_data, labels.all_data

  def test_by_window(self, dataset, window_size):
    """Yields (results, labels) windows, step window_size//2 (infer_decoder.py:484-504)."""
    storage = result_store.TwoResultStore(window_width=window_size,
                                          window_step=window_size // 2)
    for input_dict, output in dataset:
      infer_results = self.infer_one(input_dict, output)
      storage.add_data(np.reshape(infer_results, (infer_results.shape[0], -1)),
                       _to_numpy(input_dict['attended_speaker']))
      for r1, r2 in storage.next_window():
        yield r1, r2

  def compute_lda_model(self, d1, d2):
    """Fits the scaled LDA separating class-0 data d1 from class-1 data d2; returns d'
    (infer_decoder.py:506-532)."""
    if not isinstance(d1, np.ndarray):
      raise TypeError('Input d1 must be an numpy array, not %s.' % type(d1))
    if not isinstance(d2, np.ndarray):
      raise TypeError('Input d2 must be an numpy array, not %s.' % type(d2))
    data = np.concatenate((d1, d2), axis=0)
    labels = np.concatenate((1 * np.ones(d1.shape[0],), 2 * np.ones(d2.shape[0],)))
    self._lda = scaled_lda.ScaledLinearDiscriminantAnalysis()
    predictions = self._lda.fit_transform(data, labels)
    return calculate_dprime(predictions[labels == 1, 0], predictions[labels == 2, 0])

  def reduce_with_lda(self, d1):
    """Rotates the data so as to get the maximal separation in dimension 0
    (infer_decoder.py:534-550)."""
    if self._lda is None or self._lda.coef_array is None:
      raise ValueError('Must compute the LDA model before reducing data.')
    if not isinstance(d1, np.ndarray):
      raise TypeError('Input data must be an numpy array, not %s.' % type(d1))
    return self._lda.transform(d1)

  # ------------------------------------------------------------- whole-dataset (B200) interface
  def decode_dataset(self, dataset):
    """(result1, result2) float32 CUDA [kept rows, dims] in stream order for a whole dataset."""
    raise NotImplementedError('Must be implemented by a subclass.')

  def train_statistics(self, *datasets):
    """Global means and power from one pass over the datasets (infer_decoder.py:366-371)."""
    self.add_statistics(*[self.decode_dataset(ds) for ds in datasets])

  def add_statistics(self, *pairs):
    """Adds (result1, result2) CUDA [rows, dims] pairs to the running correlator statistics -- what
    add_data_correlator (infer_decoder.py:288-310) does minibatch by minibatch, in one kernel pass."""
    sums = None
    count = 0
    for r1, r2 in pairs:
      sums = ops.column_moments(r1, r2, sums)
      count += int(r1.shape[0])
    s = sums.cpu().numpy().astype(np.float32)   # the reference keeps float32 sums
    self._count += count
    self._sum_x = self._sum_x + s[0]
    self._sum_y = self._sum_y + s[1]
    self._sum_x2 = self._sum_x2 + s[2]
    self._sum_y2 = self._sum_y2 + s[3]
    self._finish_statistics()

  def train(self, data0, data1, window_size=0):
    """Trains the classification part of the decoder: global statistics, then the scaled LDA on
    window-averaged correlations; data0 = class 0 (unattended / mixed-up), data1 = class 1.
    Returns d' (infer_decoder.py:330-400).  Two passes over the data, both on the GPU."""
    for d in (data0, data1):
      if not brain_model.is_dataset(d):
        raise TypeError('Must feed training routine data0 with a tf.data.Dataset'
                        ' not a %s.' % type(d))
    self.train_statistics(data0, data1)
    averaged = []
    for ds in (data0, data1):
      r1, r2 = self.decode_dataset(ds)
      n, dims = int(r1.shape[0]), int(r1.shape[1])
      if n == 0:
        raise ValueError('No data for class %d' % len(averaged))
      mu_a, mu_b, power = self._device_statistics(dims, r1.device)
      w = max(int(window_size), 1)             # average_data: 0 or 1 means no averaging
      vals, _ = ops.windowed_corr(r1, r2, [0], [n], [w], [w], 'global_all', mu_a=mu_a, mu_b=mu_b,
                                  power=power)
      averaged.append(vals.cpu().numpy())
    return self.compute_lda_model(averaged[0], averaged[1])

  def _device_statistics(self, dims, device):
    f = lambda v: torch.as_tensor(np.broadcast_to(np.asarray(v, dtype=np.float32), (dims,)).copy(),
                                  device=device)
    return f(self._mean_x), f(self._mean_y), f(self._power)

  def windowed_results(self, dataset, window_sizes, window_steps=None):
    """{window_size: (results, labels)} float64 numpy vectors over the dataset's stream.

    Equivalent to infer.regress_and_correlate (infer.py:247-266) for every window size, but the
    model runs once and the signals are read once.  Labels are window means of the
    attended_speaker stream."""
    if self._reduction not in ops.REDUCTIONS:
      raise NotImplementedError('windowed_results supports the first / second / mean / '
                                'mean-squared reducers, not %s.' % self._reduction)
    window_sizes = [int(w) for w in window_sizes]
    if window_steps is None:
      window_steps = [w // 2 for w in window_sizes]      # infer_decoder.py:498-499
    r1, r2 = self.decode_dataset(dataset)
    n, dims = int(r1.shape[0]), int(r1.shape[1])
    mu_a, mu_b, power = self._device_statistics(dims, r1.device)
    lda_params = None
    if self._reduction == 'lda':
      if self._lda is None:
        raise ValueError('Must compute the LDA model before reducing data.')
      lda_params = torch.as_tensor(self._lda.kernel_parameters(), device=r1.device)
    vals, off = ops.windowed_corr(r1, r2, [0], [n], window_sizes, window_steps, 'global',
                                  reduction=self._reduction, mu_a=mu_a, mu_b=mu_b, power=power,
                                  lda_params=lda_params)
    labels = brain_model._stream_order(dataset, dataset.unlagged_rows('a')).contiguous()   # pylint: disable=protected-access
    labs, _ = ops.windowed_corr(labels, None, [0], [n], window_sizes, window_steps, 'plain')
    vals, labs = vals.cpu().numpy()[:, 0], labs.cpu().numpy()[:, 0]
    return {w: (vals[off[i]:off[i + 1]], labs[off[i]:off[i + 1]])
            for i, w in enumerate(window_sizes)}


class LinearRegressionDecoder(Decoder):
  """Ground truth vs. the linear model's prediction (infer_decoder.py:583-605)."""

  def decode_one(self, input_dict, ground_truth):
    predictions = self._decoding_model(input_dict)
    return _to_numpy(ground_truth), _to_numpy(predictions)

  def decode_dataset(self, dataset):
    pred = brain_model._stream_order(dataset, self._decoding_model._forward_dataset(dataset))   # pylint: disable=protected-access
    truth = brain_model._stream_order(dataset, dataset.unlagged_rows('y'))   # pylint: disable=protected-access
    return truth.contiguous(), pred.contiguous()


class CCADecoder(Decoder):
  """The two halves of the CCA output (infer_decoder.py:608-633)."""

  def decode_one(self, input_dict, ground_truth):
    del ground_truth
    predictions = _to_numpy(self._decoding_model(input_dict))
    num_cca_dims = predictions.shape[1] // 2
    return predictions[:, :num_cca_dims], predictions[:, num_cca_dims:]

  def decode_dataset(self, dataset):
    out = brain_model._stream_order(dataset, self._decoding_model._forward_dataset(dataset))   # pylint: disable=protected-access
    k = out.shape[1] // 2
    return out[:, :k].contiguous(), out[:, k:].contiguous()


def create_decoder(model_tag, reduction='lda', model=None):
  """LinearRegressionDecoder or CCADecoder from a tag (infer_decoder.py:636-667)."""
  if 'linear' in model_tag.lower() or 'fullyconnected' in model_tag.lower():
    return LinearRegressionDecoder(model, reduction=reduction)
  elif 'cca' in model_tag.lower():
    return CCADecoder(model, reduction=reduction)
  raise ValueError('Couldn\'t determine model type for tag %s.' % model_tag)


def calculate_dprime(d1, d2):
  """d' of two sets of scalars (infer_decoder.py:717-745)."""
  d1, d2 = np.asarray(d1), np.asarray(d2)
  if d1.ndim > 2 or (d1.ndim == 2 and d1.shape[1] > 1):
    raise TypeError('d1 array must be a vector, not size %s.' % str(d1.shape))
  if d2.ndim > 2 or (d2.ndim == 2 and d2.shape[1] > 1):
    raise TypeError('d2 array must be a vector, not size %s.' % str(d2.shape))
  return (np.mean(d2) - np.mean(d1)) / np.sqrt((np.var(d1) + np.var(d2)) / 2.0)


def average_data(data, window_size):
  """Non-overlapping window means, tail dropped (infer_decoder.py:748-783)."""
  if not isinstance(data, np.ndarray):
    raise TypeError('Data to be averaged must be a numpy array, not %s.' % type(data))
  if data.ndim != 2:
    raise TypeError('Averaging data must be two dimensional, not %s.' % data.ndim)
  if not window_size >= 0:
    raise ValueError('Window size (%s) must be greater-than or equal to '
                     'zero.' % window_size)
  if window_size <= 1:
    return data
  num_frames = data.shape[0] // window_size
  trimmed = data[0:num_frames * window_size, :]
  return trimmed.reshape(num_frames, window_size, data.shape[1]).mean(axis=1)
