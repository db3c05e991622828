"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from oracle import ref_numpy as ora
from telluride_decoding_b200 import brain_data


def rel_fro(a, b):
  a = np.asarray(a, dtype=np.float64)
  b = np.asarray(b, dtype=np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def make_test_data(x, y, x2=None, a=None, **kwargs):
  kwargs.setdefault('shuffle_buffer_size', 0)
  bd = brain_data.TestBrainData('input_1', 'output', 100.0, in2_fields='input_2', **kwargs)
  bd.preserve_test_data(x, y, x2, a)
  return bd


def oracle_batches(x, y, x2=None, a=None, pre=0, post=0, pre2=0, post2=0, offset=0, batch=100,
                   chunk=1000000):
  n = x.shape[0]
  x2 = np.zeros((n, 1), np.float32) if x2 is None else x2
  a = np.zeros((n, 1), np.float32) if a is None else a
  rows = ora.file_rows(x, x2, y, a, pre, post, pre2, post2, offset, chunk)
  return ora.batches_from_files([rows], batch)


def design64(batches, key='input_1'):
  return np.concatenate([np.asarray(b[0][key], dtype=np.float64) for b in batches], axis=0)
