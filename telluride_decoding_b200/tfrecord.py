"""TFRecord files of tf.train.Examples, one frame per record, without TensorFlow.

The reference reads its experiments with tf.data.TFRecordDataset + tf.io.parse_example
(telluride_decoding/brain_data.py:733-745, :758-812) and writes them with ingest.py:1118-1172.
Here the native parser of libtdgram.so (csrc/tfrecord.cu, C ABI in include/tdgram.h) turns a
whole file into one pinned [frames, width] float32 array that is a single copy away from the GPU.
"""
import ctypes
import os
import struct

import numpy as np
import torch

from telluride_decoding_b200 import _lib

KIND_NAMES = {1: 'string', 2: 'float32', 3: 'int64'}


def _file_bytes(path):
  if not isinstance(path, str):
    raise TypeError('tfrecord_file_name must be a string.')
  data = np.fromfile(path, dtype=np.uint8)
  return data, data.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(data.size)


def count_tfrecords(tfrecord_file_name, verify_crc=2):
  """(number of valid records, whether an error was found)  -- brain_data.py:895-920."""
  if not isinstance(tfrecord_file_name, str):
    raise TypeError('tfrecord_file_name must be a string.')
  lib = _lib.load()
  data, ptr, size = _file_bytes(tfrecord_file_name)
  n = ctypes.c_int64(0)
  status = lib.td_tfrecord_count(ptr, size, int(verify_crc), ctypes.byref(n))
  del data
  return int(n.value), status != 0


def feature_table(tfrecord_file_name):
  """[(name, width, kind)] of the first record's features, in file order."""
  lib = _lib.load()
  data, ptr, size = _file_bytes(tfrecord_file_name)
  max_features = 256
  names = ctypes.create_string_buffer(1 << 16)
  widths = (ctypes.c_int32 * max_features)()
  kinds = (ctypes.c_int32 * max_features)()
  n = ctypes.c_int32(0)
  _lib.check(lib.td_tfrecord_features(ptr, size, names, ctypes.c_size_t(len(names)), widths, kinds,
                                      max_features, ctypes.byref(n)))
  out, raw = [], names.raw
  pos = 0
  for i in range(n.value):
    end = raw.index(b'\0', pos)
    out.append((raw[pos:end].decode('utf-8'), int(widths[i]), int(kinds[i])))
    pos = end + 1
  return out


def read_features(tfrecord_file_name, names, widths=None, verify_crc=1, pin=None):
  """The named float features of every record, side by side: float32 [frames, sum(widths)].

  The result lives in pinned host memory when CUDA is present (pin=None) so that `.to('cuda',
  non_blocking=True)` is one asynchronous copy."""
  lib = _lib.load()
  names = [names,] if isinstance(names, str) else list(names)
  if widths is None:
    table = {n: w for n, w, _ in feature_table(tfrecord_file_name)}
    missing = [n for n in names if n not in table]
    if missing:
      raise ValueError('Could not find all desired features (%s) in data (%s)' %
                       (names, sorted(table.keys())))
    widths = [table[n] for n in names]
  data, ptr, size = _file_bytes(tfrecord_file_name)
  count = ctypes.c_int64(0)
  _lib.check(lib.td_tfrecord_count(ptr, size, 0, ctypes.byref(count)))
  total = int(sum(widths))
  if pin is None:
    pin = torch.cuda.is_available()
  out = torch.empty((int(count.value), total), dtype=torch.float32, pin_memory=bool(pin))
  c_names = (ctypes.c_char_p * len(names))(*[n.encode('utf-8') for n in names])
  c_widths = (ctypes.c_int32 * len(names))(*[int(w) for w in widths])
  got = ctypes.c_int64(0)
  _lib.check(lib.td_tfrecord_read_f32(ptr, size, c_names, c_widths, len(names),
                                      ctypes.c_void_p(out.data_ptr()), ctypes.c_int64(total),
                                      ctypes.c_int64(-1), int(verify_crc), ctypes.byref(got)))
  assert got.value == count.value
  return out


# --------------------------------------------------------------------------------------------
# Writer (tests, fixtures, exporting synthetic experiments): the format of ingest.py:1118-1172.
# --------------------------------------------------------------------------------------------

def _crc32c_table():
  table = []
  for i in range(256):
    c = i
    for _ in range(8):
      c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    table.append(c)
  return table


_CRC_TABLE = _crc32c_table()


def _masked_crc(data):
  c = 0xFFFFFFFF
  for b in data:
    c = (c >> 8) ^ _CRC_TABLE[(c ^ b) & 0xFF]
  c ^= 0xFFFFFFFF
  return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(v):
  out = bytearray()
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _len_field(field, payload):
  return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features):
  """Serialized tf.train.Example from {name: 1-D float array} (float_list features)."""
  body = b''
  for name in sorted(features.keys()):
    values = np.asarray(features[name], dtype='<f4').reshape(-1)
    float_list = _len_field(1, values.tobytes())
    feature = _len_field(2, float_list)
    entry = _len_field(1, name.encode('utf-8')) + _len_field(2, feature)
    body += _len_field(1, entry)
  return _len_field(1, body)


def write_tfrecord(path, columns):
  """One record per frame from {name: [frames, width] array}."""
  names = list(columns.keys())
  arrays = [np.asarray(columns[n], dtype=np.float32) for n in names]
  arrays = [a.reshape(a.shape[0], -1) for a in arrays]
  frames = arrays[0].shape[0]
  if any(a.shape[0] != frames for a in arrays):
    raise ValueError('all features need the same number of frames')
  os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
  with open(path, 'wb') as f:
    for t in range(frames):
      payload = encode_example({n: a[t] for n, a in zip(names, arrays)})
      header = struct.pack('<Q', len(payload))
      f.write(header + struct.pack('<I', _masked_crc(header)) + payload +
              struct.pack('<I', _masked_crc(payload)))
