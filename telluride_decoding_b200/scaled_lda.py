"""Two-class linear discriminant analysis with 0/1 output scaling (telluride_decoding/scaled_lda.py).

The reducer that turns the multi-dimensional correlation signal of a CCA decoder into one scalar
(`reduction='lda'`, infer_decoder.py:451-452).  The matrices are dims x dims with dims <= ~10, so the
fit stays on the host in float64 numpy (the reference sums one outer product per row in a Python
loop, scaled_lda.py:143-148; here the same scatter matrices come from two matrix products).  The
transform of whole recordings runs inside the windowed-correlation kernel (reduction 5).
"""

import collections

import numpy as np

LdaParamsTuple = collections.namedtuple(
    'LdaParamsTuple', ['w_real', 'w_imag', 'labels', 'mean_vectors', 'slope', 'intercept'])


class LinearDiscriminantAnalysis(object):
  """Best linear mapping to discriminate classes of vector data (scaled_lda.py:35-267)."""

  def __init__(self):
    self._eigen_pairs = []
    self._labels = []
    self._mean_vectors = []
    self._w = None

  @property
  def mean_vectors(self):
    return self._mean_vectors

  @property
  def coef_array(self):
    return self._w

  @property
  def labels(self):
    return self._labels

  @property
  def model_parameters(self):
    return LdaParamsTuple(np.real(self._w), np.imag(self._w), self._labels, self._mean_vectors,
                          None, None)

  @model_parameters.setter
  def model_parameters(self, values):
    self._set_parameters(values)

  def _set_parameters(self, values):
    if values.w_real is not None:
      self._w = np.array(values.w_real) + 1j * np.array(values.w_imag)
    else:
      self._w = None
    self._labels = np.array(values.labels)
    self._mean_vectors = np.array(values.mean_vectors)

  @classmethod
  def from_fitted_data(cls, x, y):
    obj = cls()
    obj.fit(x, y)
    return obj

  def expand_dims(self, data):
    if data.ndim == 1:
      data = np.reshape(data, (-1, 1))
    return data

  def _scatter_matrices(self, x, y):
    """(within, between) class scatter, scaled_lda.py:129-175."""
    num_dims = x.shape[1]
    within = np.zeros((num_dims, num_dims))
    between = np.zeros((num_dims, num_dims))
    overall_mean = np.mean(x, axis=0)
    for label, mean_vector in zip(self._labels, self._mean_vectors):
      rows = x[y == label]
      centred = rows - mean_vector
      within += centred.T.dot(centred)
      delta = (mean_vector - overall_mean).reshape(num_dims, 1)
      between += rows.shape[0] * delta.dot(delta.T)
    return within, between

  def fit(self, x, y):
    """Stores the rotation of the two leading discriminant directions (scaled_lda.py:177-214)."""
    x = self.expand_dims(np.asarray(x))
    y = np.asarray(y)
    self._labels = sorted(set(y))
    self._mean_vectors = [np.mean(x[y == label], axis=0) for label in self._labels]
    within, between = self._scatter_matrices(x, y)
    eigen_vals, eigen_vecs = np.linalg.eig(np.linalg.inv(within).dot(between))
    pairs = [(np.abs(eigen_vals[i]), eigen_vecs[:, i]) for i in range(len(eigen_vals))]
    self._eigen_pairs = sorted(pairs, key=lambda k: k[0], reverse=True)
    num_dims = x.shape[1]
    if len(self._eigen_pairs) > 1:
      self._w = np.hstack((self._eigen_pairs[0][1].reshape(num_dims, 1),
                           self._eigen_pairs[1][1].reshape(num_dims, 1)))
    else:
      self._w = np.array([[1,],])

  def transform(self, x):
    if self._w is None:
      raise ValueError('Must fit the model before transforming.')
    x = self.expand_dims(np.asarray(x))
    if np.ndim(x) != 2 or self._w.shape[0] != x.shape[1]:
      raise TypeError('Inconsistent training and transform sizes. %s vs %s' %
                      (x.shape, self._w.shape))
    return np.real(x.dot(self._w))

  def fit_transform(self, x, y):
    self.fit(x, y)
    return self.transform(x)

  def explained_variance_ratio(self):
    if self._w is None:
      raise ValueError('Must fit the model before transforming.')
    eigen_vals = np.array([val for val, _ in self._eigen_pairs])
    return eigen_vals / np.sum(eigen_vals)


class ScaledLinearDiscriminantAnalysis(LinearDiscriminantAnalysis):
  """LDA whose first axis maps the smaller label's mean to 0 and the larger one's to 1
  (scaled_lda.py:270-355)."""

  def __init__(self):
    self._slope = 1
    self._intercept = 0
    super(ScaledLinearDiscriminantAnalysis, self).__init__()

  @property
  def model_parameters(self):
    return LdaParamsTuple(np.real(self._w), np.imag(self._w), self._labels, self._mean_vectors,
                          self._slope, self._intercept)

  @model_parameters.setter
  def model_parameters(self, values):
    self._set_parameters(values)

  def _set_parameters(self, values):
    values = LdaParamsTuple(*values)
    super(ScaledLinearDiscriminantAnalysis, self)._set_parameters(values)
    self._slope = values.slope
    self._intercept = values.intercept

  def fit(self, x, y, y0=0, y1=1):
    x = self.expand_dims(np.asarray(x))
    self._slope, self._intercept = 1, 0
    super(ScaledLinearDiscriminantAnalysis, self).fit(x, y)
    if len(self.labels) != 2:
      raise ValueError('Scaled LDA can only be done on two-class data.')
    x0 = self.transform(np.reshape(self.mean_vectors[0], (1, -1)))[0, 0]
    x1 = self.transform(np.reshape(self.mean_vectors[1], (1, -1)))[0, 0]
    if x0 == x1:
      raise ValueError('X0 and X1 in Scaled LDA are identical (%g and %g)' % (x0, x1))
    self._slope = (y0 - y1) / (x0 - x1)
    self._intercept = y0 - self._slope * x0

  def fit_two_classes(self, class0, class1):
    class0 = np.asarray(class0)
    class1 = np.asarray(class1)
    if class0.ndim * class1.ndim != 1 and class0.shape[1] != class1.shape[1]:
      raise ValueError('Class 0 and Class1 must have the same number of '
                       'dimensions (%s vs %s).' % (class0.shape, class1.shape))
    x = np.concatenate((class0, class1), axis=0)
    y = np.concatenate((np.ones(class0.shape[0]) * 0, np.ones(class0.shape[0]) * 1))
    self.fit(x, y)

  def transform(self, x):
    x_lda = super(ScaledLinearDiscriminantAnalysis, self).transform(x)
    return np.real(self._slope * x_lda + self._intercept)

  def kernel_parameters(self):
    """float64 [w_0 .. w_{dims-1}, slope, intercept] of the first axis, for ops.windowed_corr."""
    if self._w is None:
      raise ValueError('Must compute the LDA model before reducing data.')
    return np.concatenate((np.real(self._w[:, 0]).astype(np.float64),
                           [float(self._slope), float(self._intercept)]))
