"""ctypes binding of libtdgram.so (include/tdgram.h) -- the only door to the CUDA kernels.

Host code is Python; torch is used for device memory, streams and DLPack hand-off only.
There is NO CPU fallback: if the shared library is missing or a kernel is asked to run
without a CUDA device, this module raises.
"""

import ctypes
import os

import torch
import torch.utils.dlpack

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtdgram.so')


class TdError(RuntimeError):
  """A non-zero td_status from libtdgram.so."""

  def __init__(self, status, message):
    super().__init__('libtdgram status %d: %s' % (status, message))
    self.status = status
    self.message = message


class TdStream(ctypes.Structure):
  _fields_ = [('data_dev', ctypes.c_void_p), ('ld', ctypes.c_int64),
              ('channels', ctypes.c_int32), ('pre', ctypes.c_int32), ('post', ctypes.c_int32),
              ('seg_start', ctypes.POINTER(ctypes.c_int64)),
              ('seg_len', ctypes.POINTER(ctypes.c_int64))]


class TdGramOptions(ctypes.Structure):
  _fields_ = [('method', ctypes.c_int32), ('cta_group', ctypes.c_int32),
              ('chunk_rows', ctypes.c_int32), ('superchunk_rows', ctypes.c_int32),
              ('max_ctas', ctypes.c_int32), ('reserved', ctypes.c_int32 * 3)]


# Every symbol include/tdgram.h declares (tests check the library exports all of them).
EXPORTS = (
    'td_last_error', 'td_version', 'td_device_info', 'td_lagged_gram_workspace_bytes',
    'td_lagged_gram_f32', 'td_last_gram_kernel_ms', 'td_lagged_gram_range_f32', 'td_lagged_thin_f64', 'td_lagged_predict_f32', 'td_lagged_rows_f32',
    'td_ridge_solve', 'td_cca_solve', 'td_column_moments_f64', 'td_windowed_corr_f32',
    'td_window_count', 'td_wta_decide_u8', 'td_symmetrize_f32', 'td_tfrecord_count',
    'td_tfrecord_features', 'td_tfrecord_read_f32', 'td_shutdown', 'td_comm_unique_id', 'td_comm_init',
    'td_comm_destroy', 'td_comm_nccl_version', 'td_allreduce_sum_f32', 'td_allreduce_sum_f64',
    'td_allreduce_sym_f32', 'td_pack_upper_f32', 'td_unpack_upper_f32', 'td_lagged_gram_batched_f32',
    'td_lagged_thin_per_segment_f64', 'td_lagged_predict_per_segment_f32', 'td_ridge_solve_batched',
    'td_lagged_column_sums_f64', 'td_lagged_lw_stat_f64', 'td_ridge_loo_solve', 'td_lagged_gram_batched_workspace_bytes', 'td_sum_batches_f32')

_lib = None


def load():
  """Loads libtdgram.so (once).  Raises if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(
        'libtdgram.so is missing at %s: build it with `python -c "import __graft_entry__ as g; '
        'g.build()"` or `make -C telluride_decoding_b200/csrc`. There is no CPU fallback.' %
        LIB_PATH)
  lib = ctypes.CDLL(LIB_PATH)
  lib.td_last_error.restype = ctypes.c_char_p
  lib.td_window_count.restype = ctypes.c_int64
  lib.td_window_count.argtypes = [ctypes.c_int64] * 3
  _lib = lib
  return lib


def check(status):
  if status != 0:
    raise TdError(status, load().td_last_error().decode('utf-8', 'replace'))


_cuda_checked = False


def require_cuda():
  """Raises without a CUDA device.  (torch.cuda.is_available() costs ~0.3 ms per call -- it asks
  the driver every time -- so a positive answer is remembered.)"""
  global _cuda_checked
  if _cuda_checked:
    return
  if not torch.cuda.is_available():
    raise RuntimeError('telluride_decoding_b200 needs a CUDA device (B200, sm_100a); '
                       'there is no CPU fallback.')
  _cuda_checked = True


def current_stream_ptr():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def i64_array(values):
  values = [int(v) for v in values]
  return (ctypes.c_int64 * max(len(values), 1))(*values)


def i32_array(values):
  values = [int(v) for v in values]
  return (ctypes.c_int32 * max(len(values), 1))(*values)


def f64_array(values):
  values = [float(v) for v in values]
  return (ctypes.c_double * max(len(values), 1))(*values)


def ptr(t):
  """Device pointer of a torch CUDA tensor (None -> NULL)."""
  if t is None:
    return ctypes.c_void_p(0)
  if not t.is_cuda:
    raise ValueError('expected a CUDA tensor')
  return ctypes.c_void_p(t.data_ptr())


def _is_dlpack_capsule(obj):
  return type(obj).__name__ == 'PyCapsule'


def as_device_f32(array, device=None):
  """Anything array-like -> contiguous CUDA float32 tensor.

  torch tensors are taken as they are; any other object that speaks the DLPack protocol
  (`__dlpack__` / `__dlpack_device__`: cupy, jax, another framework's GPU array) and raw DLPack
  capsules cross without a copy through torch.from_dlpack; numpy arrays and everything else go
  through numpy (host memory, one upload)."""
  require_cuda()
  import numpy as np
  if isinstance(array, torch.Tensor):
    t = array
  elif _is_dlpack_capsule(array):
    t = torch.utils.dlpack.from_dlpack(array)
  elif hasattr(array, '__dlpack__') and not isinstance(array, np.ndarray):
    t = torch.from_dlpack(array)
  else:
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(array, dtype=np.float32)))
  t = t.to(device=device or 'cuda', dtype=torch.float32)
  return t.contiguous()


class Stream(object):
  """A raw [rows, channels] CUDA tensor + lag spec + segment table -> td_stream."""

  def __init__(self, data, pre, post, seg_start, seg_len):
    if data.dim() != 2:
      raise ValueError('stream data must be 2-D (frames x channels)')
    self.data = data
    self.pre = int(pre)
    self.post = int(post)
    self.seg_start = [int(v) for v in seg_start]
    self.seg_len = [int(v) for v in seg_len]
    self._c_start = i64_array(self.seg_start)
    self._c_len = i64_array(self.seg_len)

  @property
  def channels(self):
    return int(self.data.shape[1])

  @property
  def lags(self):
    return self.pre + 1 + self.post

  @property
  def width(self):
    return self.channels * self.lags

  def c_struct(self):
    return TdStream(self.data.data_ptr(), int(self.data.stride(0)), self.channels, self.pre,
                    self.post, self._c_start, self._c_len)
