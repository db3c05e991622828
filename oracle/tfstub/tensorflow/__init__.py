"""Permissive stand-in for the `tensorflow` package.  TEST INFRASTRUCTURE ONLY.

TensorFlow is not installed in this image, but the reference's hot-path arithmetic
(`brain_model.calculate_linear_regressor_parameters_from_dataset`,
`cca.calculate_cca_parameters_from_dataset`, `infer_decoder.Decoder`'s correlator,
`result_store`, `attention_decoder`) is plain numpy.  With this stub first on
`sys.path`, `/root/reference/telluride_decoding/*.py` import UNMODIFIED and those
functions run as written.  `oracle/gen_golden.py` uses that (in the build container
only -- `/root/reference` does not travel) to pin the numpy restatement in
`oracle/ref_numpy.py` and to write the fixtures under `tests/golden/`.

Nothing here computes anything: `tf.data.Dataset` is a list of batches whose
elements expose `.numpy()` / `.shape`, `tf.function` is the identity decorator and
every other attribute resolves to a dummy class usable as a base class.
"""
import sys
import types

import numpy as np


class _Tensor(object):
  """A batch element: numpy array behind `.numpy()` / `.shape`."""

  def __init__(self, array):
    self._a = np.asarray(array)

  def numpy(self):
    return self._a

  @property
  def shape(self):
    return self._a.shape

  @property
  def dtype(self):
    return self._a.dtype

  def __getitem__(self, idx):
    return _Tensor(self._a[idx])

  def __array__(self, dtype=None, copy=None):
    return self._a if dtype is None else self._a.astype(dtype)


def _wrap(item):
  if isinstance(item, dict):
    return {k: _wrap(v) for k, v in item.items()}
  if isinstance(item, (tuple, list)):
    return tuple(_wrap(v) for v in item)
  if isinstance(item, _Tensor):
    return item
  return _Tensor(item)


class Dataset(object):
  """List-backed `tf.data.Dataset`: iteration and `take(n)` (n < 0 => all)."""

  def __init__(self, batches):
    self._batches = [_wrap(b) for b in batches]

  def __iter__(self):
    return iter(self._batches)

  def take(self, n):
    if n is None or n < 0:
      return Dataset(self._batches)
    return Dataset(self._batches[:n])

  def __len__(self):
    return len(self._batches)


class _Dummy(object):
  """Callable, subclassable, attribute-bottomless placeholder."""

  def __init__(self, *args, **kwargs):
    pass

  def __call__(self, *args, **kwargs):
    return _Dummy()

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return _Dummy()


class _Module(types.ModuleType):

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    sub = _Module(self.__name__ + '.' + name)
    # Anything CamelCase is used as a class (often a base class); the rest as modules/functions.
    if name[:1].isupper():
      cls = type(name, (_Dummy,), {})
      setattr(self, name, cls)
      return cls
    setattr(self, name, sub)
    return sub

  def __call__(self, *args, **kwargs):
    return _Dummy()


def _function(fn=None, **unused_kwargs):
  if fn is None:
    return lambda f: f
  return fn


_this = sys.modules[__name__]
_root = _Module('tensorflow')
_root.__dict__.update({k: v for k, v in _this.__dict__.items()
                       if k.startswith('__') and k not in ('__class__', '__dict__')})
_root.function = _function
_root.Tensor = _Tensor
_root.float32 = np.float32
_root.float64 = np.float64
_root.int64 = np.int64
_data = _Module('tensorflow.data')
_data.Dataset = Dataset
_data.DatasetV2 = Dataset
_root.data = _data
_root.stub_tensor = _Tensor
_root.stub_dataset = Dataset
_compat = _Module('tensorflow.compat')
_compat.v2 = _root
_compat.v1 = _root
_root.compat = _compat
sys.modules['tensorflow'] = _root
sys.modules['tensorflow.compat'] = _compat
sys.modules['tensorflow.compat.v2'] = _root
sys.modules['tensorflow.compat.v1'] = _root
