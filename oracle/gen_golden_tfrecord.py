"""Generate the TFRecord fixtures from the reference's own MEG test data.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_tfrecord.py        (build container: /root/reference is mounted)

Writes
  tests/golden/meg_subj01_head.tfrecords   the first 300 records of
      /root/reference/test_data/meg/subj01_1ksamples.tfrecords, byte for byte (the file the
      reference's test/brain_data_test.py:509-526 reads), and
  tests/golden/meg_subj01_head.npz         the same records decoded by the independent pure-Python
      protobuf walker below, plus the ridge solution of the UNMODIFIED reference function
      (brain_model.calculate_linear_regressor_parameters_from_dataset under oracle/tfstub) for
      mel_spectrogram (64 channels, 1 frame before / 2 after) -> envelope, batch 50.
"""
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'tfstub'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np                                            # noqa: E402
import tensorflow as tf                                       # noqa: E402  (the stub)
from telluride_decoding import brain_model                    # noqa: E402
from oracle import ref_numpy as ora                           # noqa: E402

SRC = '/root/reference/test_data/meg/subj01_1ksamples.tfrecords'
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
HEAD = 300


def _varint(buf, pos):
  value, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    value |= (b & 0x7F) << shift
    if not b & 0x80:
      return value, pos
    shift += 7


def _fields(buf):
  pos = 0
  while pos < len(buf):
    key, pos = _varint(buf, pos)
    wire = key & 7
    if wire == 2:
      length, pos = _varint(buf, pos)
      yield key >> 3, buf[pos:pos + length]
      pos += length
    elif wire == 0:
      _, pos = _varint(buf, pos)
    elif wire == 5:
      pos += 4
    elif wire == 1:
      pos += 8
    else:
      raise ValueError('wire type %d' % wire)


def parse_example(payload):
  """{name: float32 array} of a serialized tf.train.Example (float_list features)."""
  out = {}
  for f, features in _fields(payload):
    if f != 1:
      continue
    for f2, entry in _fields(features):
      if f2 != 1:
        continue
      parts = dict(_fields(entry))
      name = parts[1].decode('utf-8')
      for kind, lst in _fields(parts[2]):
        if kind == 2:
          for f3, packed in _fields(lst):
            if f3 == 1:
              out[name] = np.frombuffer(packed, dtype='<f4').copy()
  return out


def main():
  raw = open(SRC, 'rb').read()
  pos, records = 0, []
  while pos < len(raw) and len(records) < HEAD:
    (length,) = struct.unpack_from('<Q', raw, pos)
    records.append(raw[pos + 12:pos + 12 + length])
    pos += 16 + length
  with open(os.path.join(OUT, 'meg_subj01_head.tfrecords'), 'wb') as f:
    f.write(raw[:pos])
  cols = {}
  for rec in records:
    for k, v in parse_example(rec).items():
      cols.setdefault(k, []).append(v)
  cols = {k: np.stack(v) for k, v in cols.items()}
  x, y = cols['mel_spectrogram'], cols['envelope']
  pre, post, batch = 1, 2, 50
  zeros = np.zeros((x.shape[0], 1), np.float32)
  rows = ora.file_rows(x, zeros, y, zeros, pre, post)
  batches = ora.batches_from_files([rows], batch)
  w, b, _, _, _ = brain_model.calculate_linear_regressor_parameters_from_dataset(
      tf.data.Dataset(batches), lamb=0.1)
  np.savez_compressed(os.path.join(OUT, 'meg_subj01_head.npz'), pre=pre, post=post, batch=batch,
                      ridge_w=w, ridge_b=np.asarray(b), names=np.array(sorted(cols.keys())),
                      **{'col_' + k: v for k, v in cols.items()})
  print('wrote', HEAD, 'records,', pos, 'bytes;', {k: v.shape for k, v in cols.items()}, 'w', w.shape)


if __name__ == '__main__':
  main()
