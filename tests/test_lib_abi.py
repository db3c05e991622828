"""CPU: the C-ABI library loads, exports every symbol include/tdgram.h declares, and its
host-side argument checking / planning works without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest

from telluride_decoding_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
  text = open(os.path.join(ROOT, 'include', 'tdgram.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(td_[a-z0-9_]+)\s*\(', text)))


def test_exports_match_header():
  lib = _lib.load()
  names = _header_symbols()
  assert len(names) >= 15
  for n in names:
    assert hasattr(lib, n), 'libtdgram.so does not export %s' % n
  assert sorted(_lib.EXPORTS) == names
  assert lib.td_version() == 100


def test_window_count_matches_reference_formula():
  lib = _lib.load()
  # test/result_store_test.py:217-218: (N - W)//S + 1
  assert lib.td_window_count(1000, 201, 100) == 8
  assert lib.td_window_count(100, 201, 100) == 0
  assert lib.td_window_count(7680, 128, 64) == 119
  assert lib.td_window_count(200, 200, 100) == 1


def _stream(channels, pre, post, lens):
  starts, cur = [], 0
  for n in lens:
    starts.append(cur)
    cur += n
  s = _lib.TdStream(1 << 20, channels, channels, pre, post, _lib.i64_array(starts),
                    _lib.i64_array(lens))
  return s


def test_workspace_plan_and_argument_errors():
  lib = _lib.load()
  arr = (_lib.TdStream * 1)(_stream(64, 0, 31, [7680] * 60))
  keep = _lib.i64_array([7680] * 60)
  nbytes = ctypes.c_size_t(0)
  opt = _lib.TdGramOptions(0, 0, 0, 0, 0)
  assert lib.td_lagged_gram_workspace_bytes(arr, 1, None, keep, 60, ctypes.byref(opt),
                                            ctypes.byref(nbytes)) == 0
  # default (method 6, fp16 x 3 operand built on chip): tables + the pre-split copy, one 32-bit word per sample
  assert 460800 * 64 * 4 <= nbytes.value < 460800 * 64 * 4 + (1 << 20)
  opt4 = _lib.TdGramOptions(4, 0, 0, 0, 0)
  assert lib.td_lagged_gram_workspace_bytes(arr, 1, None, keep, 60, ctypes.byref(opt4),
                                            ctypes.byref(nbytes)) == 0
  # method 4 (3xTF32 on the raw array): tables only, the recording is read where it lies
  assert nbytes.value < (1 << 20)
  opt3 = _lib.TdGramOptions(3, 0, 0, 0, 0)
  assert lib.td_lagged_gram_workspace_bytes(arr, 1, None, keep, 60, ctypes.byref(opt3),
                                            ctypes.byref(nbytes)) == 0
  # method 3: 4 shifted raw planes of 64 channels x ~463k staged frames
  assert 4 * 64 * 4 * 460800 < nbytes.value < 2 * 4 * 64 * 4 * 460800
  # 306 channels: rows of 1224 bytes are not 16-byte aligned -> ONE padded copy (308 floats per row)
  arr306 = (_lib.TdStream * 1)(_stream(306, 0, 99, [100000] * 2))
  keep306 = _lib.i64_array([100000] * 2)
  for o in (opt, opt4):    # (the pre-split copy of method 6 has the same 308 words per row)
    assert lib.td_lagged_gram_workspace_bytes(arr306, 1, None, keep306, 2, ctypes.byref(o),
                                              ctypes.byref(nbytes)) == 0
    assert 200000 * 308 * 4 <= nbytes.value < 200000 * 308 * 4 + (1 << 20)
  # three streams are rejected with a message
  rc = lib.td_lagged_gram_workspace_bytes(arr, 3, None, keep, 60, ctypes.byref(opt), ctypes.byref(nbytes))
  assert rc == -1 and b'n_streams' in lib.td_last_error()
  bad = (_lib.TdStream * 1)(_stream(64, -1, 0, [100]))
  rc = lib.td_lagged_gram_workspace_bytes(bad, 1, None, _lib.i64_array([100]), 1, ctypes.byref(opt),
                                          ctypes.byref(nbytes))
  assert rc == -1 and b'pre_context must be >= 0' in lib.td_last_error()
  opt1 = _lib.TdGramOptions(1, 0, 0, 0, 0)
  assert lib.td_lagged_gram_workspace_bytes(arr, 1, None, keep, 60, ctypes.byref(opt1),
                                            ctypes.byref(nbytes)) == 0 and nbytes.value == 0


def test_ops_refuse_to_run_without_cuda():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  from telluride_decoding_b200 import brain_data
  bd = brain_data.TestBrainData('in', 'out', 100.0)
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    bd.preserve_test_data([[1.0, 2.0]], [[1.0]])
