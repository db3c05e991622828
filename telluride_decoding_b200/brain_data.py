"""Data layer: the reference's BrainData surface, without ever building the lagged design.

Mirrors telluride_decoding/brain_data.py (constructor arguments :99-118, create_dataset
:219-243, TestBrainData.preserve_test_data :588-642, filter_file_names :253-323,
create_brain_dataset :959-1049).  What changes is the object create_dataset returns: a
`LaggedDataset` holding the RAW per-trial [frames, channels] tensors on the GPU, the lag spec
and the row plan (segments, kept rows, batch size).  The models' fit / evaluate / predict hand
that straight to the CUDA kernels; iterating it yields the same `(dict, y)` minibatches the
reference's tf.data pipeline produces (built by td_lagged_rows_f32) for tests and compatibility.

Row-set semantics reproduced exactly:
  * context and input_offset are applied per file and per `initial_batch_size` chunk
    (brain_data.py:485-502): each chunk is a *segment* with zero padding at both ends;
  * files are interleaved round-robin, one row at a time (brain_data.py:725-728);
  * `batch(final_batch_size, drop_remainder=True)` drops the tail of that stream
    (brain_data.py:369-370), which is a per-file prefix of kept rows.
Shuffling (shuffle_buffer_size > 0) cannot be reproduced bit-for-bit (tf.data's RNG); the
batches are then shuffled with a torch generator and the dropped remainder is taken from the
un-shuffled tail.  The reference's regression presets run with shuffle_buffer_size = 0
(regression.py:459, :479).
"""

import os
import random
import re

import numpy as np
import torch

from telluride_decoding_b200 import _lib
from telluride_decoding_b200 import ops


class TensorSpec(object):
  """Stand-in for tf.TensorSpec: `.shape[-1]` is what the models read (brain_model.py:329-330)."""

  def __init__(self, shape, dtype=torch.float32):
    self.shape = tuple(shape)
    self.dtype = dtype

  def __repr__(self):
    return 'TensorSpec(shape=%s)' % (self.shape,)


class FixedLenFeature(object):
  """Stand-in for tf.io.FixedLenFeature (brain_data.py:616-619): shape + dtype."""

  def __init__(self, shape, dtype='float32'):
    self.shape = list(shape)
    self.dtype = dtype


def interleave_kept_prefix(lengths, kept_total):
  """Rows each file keeps when the round-robin stream is cut after `kept_total` elements.

  Stream order: round r emits row r of every file with more than r rows, in file order
  (tf.data interleave, cycle_length = n_files, block_length = 1; brain_data.py:725-728).
  """
  lengths = np.asarray([int(n) for n in lengths], dtype=np.int64)
  remaining = int(kept_total)
  if remaining <= 0 or lengths.size == 0:
    return [0] * int(lengths.size)
  # Largest round count R such that sum_f min(n_f, R) <= kept_total.
  lo, hi = 0, int(lengths.max())
  while lo < hi:
    mid = (lo + hi + 1) // 2
    if int(np.minimum(lengths, mid).sum()) <= remaining:
      lo = mid
    else:
      hi = mid - 1
  full_rounds = lo
  kept = np.minimum(lengths, full_rounds)
  remaining -= int(kept.sum())
  if remaining > 0:                       # the cut falls inside round `full_rounds`: first files in order
    longer = np.nonzero(lengths > full_rounds)[0][:remaining]
    kept[longer] += 1
  return [int(k) for k in kept]


def interleave_order(lengths):
  """(file index, row index) of every element of the interleaved stream, in order."""
  lengths = np.asarray(lengths, dtype=np.int64)
  if lengths.size == 0 or lengths.max() == 0:
    return np.zeros((0,), np.int64), np.zeros((0,), np.int64)
  rounds = np.arange(int(lengths.max()), dtype=np.int64)
  alive = rounds[:, None] < lengths[None, :]            # [round, file]
  r_idx, f_idx = np.nonzero(alive)                      # row-major: round, then file order
  return f_idx.astype(np.int64), r_idx.astype(np.int64)


class LaggedDataset(object):
  """Raw trials + lag spec + row plan: the object the models consume natively.

  Iterating yields `({'input_1', 'input_2', 'attended_speaker'}, y)` float32 CUDA minibatches
  exactly as brain_data.final_shuffle_and_batch (:325-393) would.
  """

  def __init__(self, x, x2, y, a, files, pre, post, pre2, post2, input_offset,
               initial_batch_size, final_batch_size, repeat_count, shuffle_buffer_size, mode,
               temporal_context=True, offset_counts=True, mixup_batch=False):
    # x, x2, y, a: concatenated [frames, channels] CUDA tensors; files: [(name, start, length)].
    self.x, self.x2, self.y, self.a = x, x2, y, a
    self.files = list(files)
    # [(row_end, cuda event)] when the tensors are still being filled by asynchronous host-to-device
    # copies (TrialBrainData.set_trials with pinned host arrays): rows < row_end are valid once the
    # event has completed.  None = everything is resident.
    self.ready = None
    self.mode = mode
    self.final_batch_size = int(final_batch_size)
    self.repeat_count = int(repeat_count) if mode == 'train' else 1
    shuffles = mode != 'program_test' and shuffle_buffer_size > 0
    self.shuffle_buffer_size = int(shuffle_buffer_size) if shuffles else 0
    self.mixup_batch = mixup_batch
    wants = temporal_context and (pre or post or pre2 or post2 or
                                  (input_offset if offset_counts else 0))
    if wants:
      self.pre, self.post, self.pre2, self.post2 = int(pre), int(post), int(pre2), int(post2)
      self.input_offset = int(input_offset)
      chunk = int(initial_batch_size)
    else:
      self.pre = self.post = self.pre2 = self.post2 = self.input_offset = 0
      chunk = None
    # ---- segments: one per (file, chunk)                      brain_data.py:492, :466-483
    seg_file, sx, lx, s2, l2, rows = [], [], [], [], [], []
    o = self.input_offset
    for f, (_, start, length) in enumerate(self.files):
      step = chunk or max(length, 1)
      for c0 in range(0, length, step):
        n = min(step, length - c0)
        n_rows = max(n - abs(o), 0)
        seg_file.append(f)
        sx.append(start + c0 + max(o, 0)); lx.append(max(n - max(o, 0), 0))
        s2.append(start + c0 + max(-o, 0)); l2.append(max(n - max(-o, 0), 0))
        rows.append(n_rows)
    self.seg_file = seg_file
    self.seg_rows = rows
    self._sx, self._lx, self._s2, self._l2 = sx, lx, s2, l2
    # ---- kept rows: per-file prefix of the interleaved stream         brain_data.py:369-370
    file_rows = [0] * len(self.files)
    for f, n in zip(seg_file, rows):
      file_rows[f] += n
    self.file_rows = file_rows
    self.total_rows = sum(file_rows)
    b = self.final_batch_size
    stream_len = self.total_rows * self.repeat_count
    self.num_batches = stream_len // b if b > 0 else 0
    kept_stream = self.num_batches * b
    self.full_repeats = kept_stream // self.total_rows if self.total_rows else 0
    self.partial_rows = kept_stream - self.full_repeats * self.total_rows
    self.element_spec = (
        {'input_1': TensorSpec((None, x.shape[1] * (self.pre + 1 + self.post))),
         'input_2': TensorSpec((None, x2.shape[1] * (self.pre2 + 1 + self.post2))),
         'attended_speaker': TensorSpec((None, a.shape[1]))},
        TensorSpec((None, y.shape[1])))

  # ---------------------------------------------------------------- plans for the kernels
  def seg_keep_for_prefix(self, kept_total):
    """Kept rows per segment when the stream is cut after `kept_total` elements."""
    per_file = interleave_kept_prefix(self.file_rows, kept_total)
    keep, used = [], [0] * len(self.files)
    for f, n in zip(self.seg_file, self.seg_rows):
      k = min(n, max(per_file[f] - used[f], 0))
      keep.append(k)
      used[f] += n
    return keep

  def passes(self):
    """[(weight, seg_keep)]: the kept multiset of rows as weighted prefixes of one epoch."""
    out = []
    if self.full_repeats:
      out.append((float(self.full_repeats), list(self.seg_rows)))
    if self.partial_rows:
      out.append((1.0, self.seg_keep_for_prefix(self.partial_rows)))
    return out

  def batch_pieces(self):
    """[(segment, first row, end row, minibatch)]: the contiguous row range of every segment that
    falls into each minibatch of the (un-shuffled, un-repeated) interleaved stream -- a minibatch is
    `final_batch_size` consecutive stream elements, i.e. a few consecutive rows of every file
    (brain_data.py:725-728, :369-370)."""
    pieces = []
    prev = [0] * len(self.seg_rows)
    b = self.final_batch_size
    n_batches = self.total_rows // b if b > 0 else 0
    for batch in range(n_batches):
      cur = self.seg_keep_for_prefix((batch + 1) * b)
      for s, (lo, hi) in enumerate(zip(prev, cur)):
        if hi > lo:
          pieces.append((s, lo, hi, batch))
      prev = cur
    return pieces

  def wait_ready(self):
    """Makes the current stream wait for every pending host-to-device copy of the trials."""
    if self.ready:
      torch.cuda.current_stream().wait_event(self.ready[-1][1])
      self.ready = None

  def segment_groups(self):
    """[(cuda event or None, [segment indices])]: segments grouped by the copy chunk that completes
    them, in copy order, so that a sum over segments can start before the last trial has landed."""
    if not self.ready:
      return [(None, list(range(len(self.seg_rows))))]
    ends = [e for e, _ in self.ready]
    groups = [[] for _ in ends]
    for i, (sx, lx, s2, l2) in enumerate(zip(self._sx, self._lx, self._s2, self._l2)):
      last = max(sx + lx, s2 + l2)
      g = 0
      while g + 1 < len(ends) and last > ends[g]:
        g += 1
      groups[g].append(i)
    return [(ev, seg) for (_, ev), seg in zip(self.ready, groups) if seg]

  def stream(self, which, segments=None, wait=True):
    """ops.Stream of 'x' (input_1), 'x2' (input_2), 'y' (output) or 'a' (attention); optionally of
    a subset of the segments (wait=False: the caller orders itself after the copy events)."""
    if wait:
      self.wait_ready()
    if segments is not None:
      full = self.stream(which, wait=False)
      return ops.Stream(full.data, full.pre, full.post, [full.seg_start[i] for i in segments],
                        [full.seg_len[i] for i in segments])
    if which == 'x':
      return ops.Stream(self.x, self.pre, self.post, self._sx, self._lx)
    data = {'x2': self.x2, 'y': self.y, 'a': self.a}[which]
    pre, post = (self.pre2, self.post2) if which == 'x2' else (0, 0)
    if which == 'a':      # the attention stream is not shifted by input_offset (brain_data.py:471-475)
      o = self.input_offset
      starts = [s - max(o, 0) for s in self._sx]
      lens = [l + max(o, 0) for l in self._lx]
      return ops.Stream(data, 0, 0, starts, lens)
    return ops.Stream(data, pre, post, self._s2, self._l2)

  def seg_out_rows(self):
    """First output row of each segment when all rows are laid out file after file."""
    out, cur = [], 0
    for n in self.seg_rows:
      out.append(cur)
      cur += n
    return out

  def stream_permutation(self, kept_total=None):
    """int64 CUDA index: position in file-major row order of each stream element (one epoch).
    Built once per dataset (evaluate / predict ask for it on every call)."""
    full = getattr(self, '_stream_perm', None)
    if full is None:
      fi, ri = interleave_order(self.file_rows)
      file_base = np.concatenate(([0], np.cumsum(self.file_rows)[:-1])).astype(np.int64)
      full = self._stream_perm = torch.from_numpy(file_base[fi] + ri).to(self.x.device)
    return full if kept_total is None else full[:kept_total]

  def unlagged_rows(self, which):
    """Rows of an un-lagged stream ('y', 'a', or 'x2' when it has no context) in file-major order,
    [total_rows, channels]."""
    s = self.stream(which)
    if s.pre or s.post:
      raise ValueError('stream %s has temporal context; its rows are not a slice of the raw data' % which)
    parts = [s.data[st:st + n] for st, n in zip(s.seg_start, self.seg_rows)]
    return torch.cat(parts, dim=0) if len(parts) != 1 else parts[0]

  # ---------------------------------------------------------------- reference-style iteration
  def __iter__(self):
    if self.num_batches == 0:
      return
    out_rows = self.seg_out_rows()
    x = ops.lagged_rows(self.stream('x'), self.seg_rows, out_rows, self.total_rows)
    x2 = ops.lagged_rows(self.stream('x2'), self.seg_rows, out_rows, self.total_rows)
    y = self.unlagged_rows('y')
    a = self.unlagged_rows('a')
    perm = self.stream_permutation()
    if self.repeat_count > 1:
      perm = perm.repeat(self.repeat_count)
    b = self.final_batch_size
    perm = perm[:self.num_batches * b]
    if self.shuffle_buffer_size > 0:
      gen = torch.Generator(device='cpu')
      gen.manual_seed(random.getrandbits(31))
      perm = perm[torch.randperm(perm.numel(), generator=gen).to(perm.device)]
    for i in range(self.num_batches):
      sel = perm[i * b:(i + 1) * b]
      bx2, by = x2[sel], y[sel]
      if self.mixup_batch:           # brain_data.py:378-382: shuffle x2 and y within the batch
        bx2 = bx2[torch.randperm(b, device=bx2.device)]
        by = by[torch.randperm(b, device=by.device)]
      yield ({'input_1': x[sel], 'input_2': bx2, 'attended_speaker': a[sel]}, by)

  def take(self, count):
    """First `count` minibatches (count < 0: all), like tf.data.Dataset.take."""
    batches = []
    for i, item in enumerate(self):
      if 0 <= count <= i:
        break
      batches.append(item)
    return batches

  def __repr__(self):
    return ('<LaggedDataset mode=%s files=%d rows=%d batches=%d x %d ctx=(%d,%d|%d,%d) off=%d>' %
            (self.mode, len(self.files), self.total_rows, self.num_batches,
             self.final_batch_size, self.pre, self.post, self.pre2, self.post2,
             self.input_offset))


class BrainData(object):
  """Describes the data of an experiment (brain_data.py:80-199); base class."""

  def __init__(self, in_fields, out_field, frame_rate, pre_context=0, post_context=0,
               in2_fields=None, in2_pre_context=0, in2_post_context=0, input_offset=0,
               attended_field=None, initial_batch_size=1000000, final_batch_size=1000,
               repeat_count=1, shuffle_buffer_size=1000, data_dir=None, data_pattern='',
               train_file_pattern='', validate_file_pattern='', test_file_pattern=''):
    if not in_fields:
      raise ValueError('Must specify at least one input field.')
    if not out_field:
      raise ValueError('Must specify an output field.')
    if frame_rate < 0:
      raise ValueError('frame_rate must be >= 0')
    if pre_context < 0:
      raise ValueError('pre_context must be >= 0')
    if post_context < 0:
      raise ValueError('post_context must be >= 0')
    if isinstance(in_fields, str):
      in_fields = [in_fields,]
    self.in1_fields = in_fields
    if isinstance(in2_fields, str) and in2_fields:
      in2_fields = [in2_fields,]
    self.in2_fields = in2_fields
    self.out_field = out_field
    self.frame_rate = frame_rate
    self.in1_pre_context = pre_context
    self.in1_post_context = post_context
    self.in2_pre_context = in2_pre_context
    self.in2_post_context = in2_post_context
    self.input_offset = input_offset
    self.attended_field = attended_field
    self.initial_batch_size = initial_batch_size
    self.final_batch_size = final_batch_size
    self.repeat_count = repeat_count
    self.shuffle_buffer_size = shuffle_buffer_size
    self.data_dir = data_dir
    self.data_pattern = data_pattern
    self.train_file_pattern = train_file_pattern or ''
    self.validate_file_pattern = validate_file_pattern or ''
    self.test_file_pattern = test_file_pattern or ''
    self.use_saved_data = False
    self._cached_file_names = []
    self.all_files()

  def all_files(self, max_count=0):
    """List of files (trials) available; shuffled once (brain_data.py:201-210)."""
    if not self._cached_file_names:
      self._get_data_file_names()
      if self._cached_file_names:
        random.shuffle(self._cached_file_names)
    if max_count > 0 and len(self._cached_file_names) > max_count:
      return self._cached_file_names[:max_count]
    return self._cached_file_names

  def set_file_patterns(self, train, validate, test):
    self.train_file_pattern = train
    self.validate_file_pattern = validate
    self.test_file_pattern = test

  def create_dataset(self, mode='train', temporal_context=True, mixup_batch=False):
    raise NotImplementedError

  def _get_data_file_names(self):
    self._cached_file_names = []

  def filter_file_names(self, mode):
    """Files used in `mode` (brain_data.py:253-323), incl. the allbut / allbut_NN patterns."""
    if mode == 'program_test':
      mode = 'test'
    if mode not in set(['test', 'validate', 'train']):
      raise ValueError('mode must be one of test, validate or train')
    filename_list = self.all_files()
    if not isinstance(filename_list, list):
      raise TypeError('Filename_list is a %s, not a list.' % type(filename_list))
    if mode.startswith('test'):
      pattern_re = re.compile(self.test_file_pattern)
    elif mode.startswith('validate'):
      pattern_re = re.compile(self.validate_file_pattern)
    else:
      if self.train_file_pattern == 'allbut':
        pattern_re = re.compile('')
      else:
        pattern_re = re.compile(self.train_file_pattern)
    if mode == 'train' and self.train_file_pattern.startswith('allbut'):
      if not (self.test_file_pattern and self.validate_file_pattern):
        raise ValueError('Both test and validate must be specified if using '
                         'allbut pattern')
      test_re = re.compile(self.test_file_pattern)
      validate_re = re.compile(self.validate_file_pattern)
      filename_list = [f for f in filename_list
                       if not (test_re.search(f) or validate_re.search(f))]
      if self.train_file_pattern.startswith('allbut_'):
        allbut = self.train_file_pattern.replace('allbut_', '', 1)
        if allbut.isdigit():
          count = int(allbut)
        else:
          raise ValueError('allbut_ spec must be an integer, not %s.' % allbut)
        if count < len(filename_list):
          filename_list = filename_list[:count]
    else:
      filename_list = [f for f in filename_list if pattern_re.search(f)]
    return filename_list

  def input_fields_width(self, input_number=1):
    """Feature width after adding context (brain_data.py:505-541)."""
    if input_number != 1 and input_number != 2:
      raise ValueError('Only 1st or 2nd input is supported here.')
    fields = self.in1_fields if input_number == 1 else self.in2_fields
    if isinstance(fields, str) and fields:
      fields = [fields,]
    if fields:
      for k in fields:
        if k not in list(self.features.keys()):
          raise TypeError('Can\'t find **%s** in valid features: %s' %
                          (k, [','.join(list(self.features.keys()))]))
      widths = [self.features[k].shape[0] for k in fields]
    else:
      widths = [1]
    if input_number == 1:
      return sum(widths) * (self.in1_pre_context + 1 + self.in1_post_context)
    return sum(widths) * (self.in2_pre_context + 1 + self.in2_post_context)

  def output_field_width(self):
    if self.out_field not in list(self.features.keys()):
      raise ValueError('Could not find output_field **%s** in %s' %
                       (self.out_field, self.features.keys()))
    return self.features[self.out_field].shape[0]

  # Shared by the in-memory classes below.
  def _make_dataset(self, mode, files, temporal_context, mixup_batch, offset_counts):
    ds = self._new_dataset(mode, files, temporal_context, mixup_batch, offset_counts)
    ds.ready = getattr(self, '_ready', None)
    return ds

  def _new_dataset(self, mode, files, temporal_context, mixup_batch, offset_counts):
    return LaggedDataset(
        self._x, self._x2, self._y, self._a, files, self.in1_pre_context,
        self.in1_post_context, self.in2_pre_context, self.in2_post_context, self.input_offset,
        self.initial_batch_size, self.final_batch_size, self.repeat_count,
        self.shuffle_buffer_size, mode, temporal_context=temporal_context,
        offset_counts=offset_counts, mixup_batch=mixup_batch)


def _check_2d(name, data):
  if data.dim() != 2:
    raise ValueError('%s must be 2-D (frames x channels), not %s' % (name, tuple(data.shape)))
  return data


class TestBrainData(BrainData):
  """Dataset which produces fixed (saved) values (brain_data.py:550-642)."""

  __test__ = False   # not a pytest class

  def create_dataset(self, mode='train', temporal_context=True, mixup_batch=False):
    if not hasattr(self, '_x'):
      raise ValueError('Must call preserve_test_data before create_dataset.')
    files = [('saved', 0, int(self._x.shape[0]))]
    return self._make_dataset(mode, files, temporal_context, mixup_batch, offset_counts=True)

  def preserve_test_data(self, input_data, output_data, input2_data=None, attention_data=None):
    """Puts arrays (numpy, torch or DLPack) into the dataset (brain_data.py:588-642)."""
    x = _check_2d('input_data', _lib.as_device_f32(input_data))
    y = _check_2d('output_data', _lib.as_device_f32(output_data))
    if x.shape[0] != y.shape[0]:
      raise ValueError('input shape (%s) and output shape (%s) are not equal.' %
                       (tuple(x.shape), tuple(y.shape)))
    self.saved_input_data = x
    self.saved_output_data = y
    self.num_input_channels = int(x.shape[1])
    self.num_output_channels = int(y.shape[1])
    self.features = {'input_1': FixedLenFeature([int(x.shape[1]),]),
                     'output': FixedLenFeature([int(y.shape[1]),])}
    if input2_data is None:
      x2 = torch.zeros((x.shape[0], 1), dtype=torch.float32, device=x.device)
    else:
      x2 = _check_2d('input2_data', _lib.as_device_f32(input2_data))
    if x.shape[0] != x2.shape[0]:
      raise ValueError('input shape (%s) and input2 shape (%s) are not equal.' %
                       (tuple(x.shape), tuple(x2.shape)))
    self.saved_input2_data = x2
    self.features['input_2'] = FixedLenFeature([int(x2.shape[1]),])
    if attention_data is None:
      a = torch.zeros((x.shape[0], 1), dtype=torch.float32, device=x.device)
    else:
      a = _check_2d('attention_data', _lib.as_device_f32(attention_data))
    if x.shape[0] != a.shape[0]:
      raise ValueError('input shape (%s) and attention shape (%s) are not equal.' %
                       (tuple(x.shape), tuple(a.shape)))
    self.saved_attention_data = a
    self.features['attention'] = FixedLenFeature([int(a.shape[1]),])
    self._x, self._x2, self._y, self._a = x, x2, y, a


class TrialBrainData(BrainData):
  """Several named trials held in GPU memory: the in-memory equivalent of TFExampleData.

  TFExampleData (brain_data.py:645-731) reads one TFRecord file per trial, selects files by
  regular expression (train / validate / test, `allbut`), adds context per file and interleaves
  the files.  This class keeps exactly those semantics with `add_trial(name, ...)` in place of
  files on disk (the TFRecord parser itself is outside the hot path).
  Like TFExampleData, an input_offset without any context is ignored (brain_data.py:761-763).
  """

  def __init__(self, *args, **kwargs):
    self._trials = {}
    self._dirty = True
    super(TrialBrainData, self).__init__(*args, **kwargs)

  def _get_data_file_names(self):
    self._cached_file_names = list(self._trials.keys())

  def add_trial(self, name, input_data, output_data, input2_data=None, attention_data=None):
    x = _check_2d('input_data', _lib.as_device_f32(input_data))
    y = _check_2d('output_data', _lib.as_device_f32(output_data))
    n = int(x.shape[0])
    if y.shape[0] != n:
      raise ValueError('input shape (%s) and output shape (%s) are not equal.' %
                       (tuple(x.shape), tuple(y.shape)))
    x2 = (torch.zeros((n, 1), dtype=torch.float32, device=x.device) if input2_data is None
          else _check_2d('input2_data', _lib.as_device_f32(input2_data)))
    a = (torch.zeros((n, 1), dtype=torch.float32, device=x.device) if attention_data is None
         else _check_2d('attention_data', _lib.as_device_f32(attention_data)))
    if x2.shape[0] != n or a.shape[0] != n:
      raise ValueError('all streams of a trial must have %d frames' % n)
    self._trials[str(name)] = (x, x2, y, a)
    self._cached_file_names = []
    self._dirty = True
    # Feature table keyed by the experiment's own field names (the first input field carries the
    # whole input width), plus the generic names TestBrainData uses.
    self.features = {'input_1': FixedLenFeature([int(x.shape[1]),]),
                     'output': FixedLenFeature([int(y.shape[1]),]),
                     'input_2': FixedLenFeature([int(x2.shape[1]),]),
                     'attention': FixedLenFeature([int(a.shape[1]),])}
    self.features[self.in1_fields[0]] = FixedLenFeature([int(x.shape[1]),])
    for extra in self.in1_fields[1:]:
      self.features.setdefault(extra, FixedLenFeature([0,]))
    self.features[self.out_field] = FixedLenFeature([int(y.shape[1]),])
    if self.in2_fields:
      self.features[self.in2_fields[0]] = FixedLenFeature([int(x2.shape[1]),])
    if self.attended_field:
      self.features[self.attended_field] = FixedLenFeature([int(a.shape[1]),])

  def set_trials(self, names, lengths, input_data, output_data, input2_data=None,
                 attention_data=None, upload_rows=None):
    """All trials at once from arrays that already hold them back to back (no per-trial copies):
    trial i is rows [sum(lengths[:i]), sum(lengths[:i+1])) of every array.

    Pinned host tensors are copied to the GPU asynchronously in a few chunks cut at trial
    boundaries, on a side stream; the datasets made from this object know which rows each copy
    completes, so the Gram of the first trials runs while the later ones are still on the PCIe bus.
    The host arrays must not be modified until the copies have completed (first use of a dataset
    made from this object); this object keeps them alive.

    upload_rows = (row_begin, row_end): time-sharded fits (sharding.needed_row_range) -- only these
    rows of the pinned host arrays are copied, the rest of the device arrays is zero; the host
    rows outside the range are never read.  The host arrays may also hold JUST those rows
    (row_end - row_begin rows instead of sum(lengths)): a rank then pins only its own share."""
    self._ready = None
    pinned = (isinstance(input_data, torch.Tensor) and isinstance(output_data, torch.Tensor) and
              input_data.device.type == 'cpu' and output_data.device.type == 'cpu' and
              input_data.is_pinned() and output_data.is_pinned() and
              input_data.dtype == torch.float32 and output_data.dtype == torch.float32 and
              input_data.dim() == 2 and output_data.dim() == 2 and
              input_data.is_contiguous() and output_data.is_contiguous() and
              input_data.shape[0] == output_data.shape[0] and len(lengths) > 1)
    if upload_rows is not None and not pinned:
      raise ValueError('upload_rows needs pinned, contiguous float32 host tensors')
    if pinned:
      x, y = self._async_upload(input_data, output_data, lengths, upload_rows)
    else:
      x = _check_2d('input_data', _lib.as_device_f32(input_data))
      y = _check_2d('output_data', _lib.as_device_f32(output_data))
    n = int(x.shape[0])
    if sum(int(l) for l in lengths) != n or len(names) != len(lengths):
      raise ValueError('lengths must sum to the %d frames of input_data, one per name' % n)
    if y.shape[0] != n:
      raise ValueError('input shape (%s) and output shape (%s) are not equal.' %
                       (tuple(x.shape), tuple(y.shape)))
    x2 = (torch.zeros((n, 1), dtype=torch.float32, device=x.device) if input2_data is None
          else _check_2d('input2_data', _lib.as_device_f32(input2_data)))
    a = (torch.zeros((n, 1), dtype=torch.float32, device=x.device) if attention_data is None
         else _check_2d('attention_data', _lib.as_device_f32(attention_data)))
    if x2.shape[0] != n or a.shape[0] != n:
      raise ValueError('all streams must have %d frames' % n)
    self._x, self._x2, self._y, self._a = x, x2, y, a
    self._offsets, self._trials, cur = {}, {}, 0
    for name, length in zip(names, lengths):
      self._offsets[str(name)] = (cur, int(length))
      self._trials[str(name)] = None          # data lives in the packed arrays
      cur += int(length)
    self._cached_file_names = []
    self._dirty = False
    self.features = {'input_1': FixedLenFeature([int(x.shape[1]),]),
                     'output': FixedLenFeature([int(y.shape[1]),]),
                     'input_2': FixedLenFeature([int(x2.shape[1]),]),
                     'attention': FixedLenFeature([int(a.shape[1]),])}
    self.features[self.in1_fields[0]] = FixedLenFeature([int(x.shape[1]),])
    self.features[self.out_field] = FixedLenFeature([int(y.shape[1]),])
    if self.in2_fields:
      self.features[self.in2_fields[0]] = FixedLenFeature([int(x2.shape[1]),])
    if self.attended_field:
      self.features[self.attended_field] = FixedLenFeature([int(a.shape[1]),])

  _copy_streams = {}
  # Cumulative fractions of the frames at which an upload chunk ends: a small first chunk so that
  # the GPU starts early, then chunks that grow with the time the previous one takes to process
  # (the Gram needs ~3x longer per frame than the PCIe copy).
  UPLOAD_CUTS = (0.1, 0.4, 1.0)

  @classmethod
  def upload_cuts(cls, lengths):
    """Row ends of the upload chunks: the first trial boundary at or after each UPLOAD_CUTS fraction."""
    n = sum(int(l) for l in lengths)
    cuts, cur = [], 0
    for i, length in enumerate(lengths):                 # chunk ends at trial boundaries
      cur += int(length)
      if len(cuts) < len(cls.UPLOAD_CUTS) and (cur >= cls.UPLOAD_CUTS[len(cuts)] * n or
                                               i == len(lengths) - 1):
        cuts.append(cur)
    if cuts[-1] != n:
      cuts.append(n)
    return cuts

  def _async_upload(self, x_host, y_host, lengths, upload_rows=None):
    """Device copies of two pinned [frames, channels] arrays, issued chunk by chunk on a side stream;
    sets self._ready = [(row_end, event)]."""
    _lib.require_cuda()
    dev = torch.device('cuda', torch.cuda.current_device())
    alloc = torch.empty if upload_rows is None else torch.zeros
    n = sum(int(l) for l in lengths)
    # host arrays that hold only rows [row_begin, row_end): host row i is device row row_begin + i
    host_off = 0
    if upload_rows is not None and int(x_host.shape[0]) != n:
      if int(x_host.shape[0]) != int(upload_rows[1]) - int(upload_rows[0]):
        raise ValueError('host arrays must hold either all %d frames or exactly the %d rows of upload_rows' %
                         (n, int(upload_rows[1]) - int(upload_rows[0])))
      host_off = int(upload_rows[0])
    x = alloc((n, x_host.shape[1]), dtype=torch.float32, device=dev)
    y = alloc((n, y_host.shape[1]), dtype=torch.float32, device=dev)
    side = TrialBrainData._copy_streams.get(dev.index)
    if side is None:
      side = TrialBrainData._copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())        # the buffers may be recycled memory
    cuts = self.upload_cuts(lengths)
    lo, hi = (0, n) if upload_rows is None else (max(int(upload_rows[0]), 0), min(int(upload_rows[1]), n))
    ready, r0 = [], lo
    with torch.cuda.stream(side):
      for cut in cuts:
        r1 = min(max(cut, lo), hi) if cut != cuts[-1] else hi
        if r1 > r0:
          x[r0:r1].copy_(x_host[r0 - host_off:r1 - host_off], non_blocking=True)
          y[r0:r1].copy_(y_host[r0 - host_off:r1 - host_off], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
        # Rows before `lo` and after `hi` are zeros that no design row of this rank reads: the last
        # chunk's event covers everything up to the end of the arrays.
        ready.append((cut if cut != cuts[-1] else n, ev))
        r0 = max(r0, r1)
    x.record_stream(side)
    y.record_stream(side)
    self._ready = ready
    self._host_arrays = (x_host, y_host)            # keep the pinned sources alive while in flight
    return x, y

  def _pack(self):
    if not self._dirty:
      return
    if any(v is None for v in self._trials.values()):
      raise ValueError('add_trial cannot be mixed with set_trials; call set_trials again.')
    names = sorted(self._trials.keys())
    self._offsets = {}
    cur = 0
    for n in names:
      self._offsets[n] = (cur, int(self._trials[n][0].shape[0]))
      cur += self._offsets[n][1]
    cat = lambda i: torch.cat([self._trials[n][i] for n in names], dim=0)
    self._x, self._x2, self._y, self._a = cat(0), cat(1), cat(2), cat(3)
    self._dirty = False

  def create_dataset(self, mode='train', temporal_context=True, mixup_batch=False,
                     mismatch_batch=False):
    if mismatch_batch:
      raise NotImplementedError('mismatch_batch randomisation is outside the decoding hot path.')
    filename_list = self.filter_file_names(mode)
    if not filename_list:
      raise ValueError('No files to process in mode %s from directory %s: %s' %
                       (mode, self.data_dir, self.all_files()))
    self._pack()
    files = [(n,) + self._offsets[n] for n in filename_list]
    return self._make_dataset(mode, files, temporal_context, mixup_batch, offset_counts=False)


def discover_feature_shapes(tfrecord_file_name):
  """{feature name: FixedLenFeature([width], dtype)} of the first record (brain_data.py:853-892)."""
  if not isinstance(tfrecord_file_name, str):
    raise TypeError('discover_feature_shapes: input must be a string filename.')
  from telluride_decoding_b200 import tfrecord
  return {name: FixedLenFeature([width,], tfrecord.KIND_NAMES.get(kind, 'float32'))
          for name, width, kind in tfrecord.feature_table(tfrecord_file_name)}


def count_tfrecords(tfrecord_file_name):
  """(valid records, whether an error was found) of a TFRecord file (brain_data.py:895-920)."""
  if not isinstance(tfrecord_file_name, str):
    raise TypeError('tfrecord_file_name must be a string.')
  from telluride_decoding_b200 import tfrecord
  return tfrecord.count_tfrecords(tfrecord_file_name)


class TFExampleData(TrialBrainData):
  """Dataset of tf.train.Examples in several .tfrecords files, one trial per file, one frame per
  record (brain_data.py:645-839).

  File discovery, the feature table, field selection (`in1_fields` side by side, the `ones`
  output of :778-780, the input-2 and attention placeholders of :795-810) and the per-file
  context / interleave / batch semantics follow the reference; the records themselves are parsed
  by the native reader (csrc/tfrecord.cu) into pinned memory and copied to the GPU once per file,
  on first use."""

  def __init__(self, *args, **kwargs):
    self._file_data = {}
    super(TFExampleData, self).__init__(*args, **kwargs)

  def _get_data_file_names(self):
    if not self.data_dir:
      raise ValueError('Missing data_dir in TFExampleData initialization. '
                       'Must specify the source of the data (FLAGS.tfrecords).')
    if not isinstance(self.data_dir, str):
      raise TypeError('data_dir must be a string, not a %s (**%s**)' %
                      (type(self.data_dir), self.data_dir))
    self._cached_file_names = []
    for path, _, files in os.walk(self.data_dir):
      self._cached_file_names += [os.path.join(path, f) for f in sorted(files)
                                  if (f.endswith('.tfrecords') and '-bad-' not in f and
                                      self.data_pattern in f)]
    if not self._cached_file_names:
      raise ValueError('Should not have an empty list of data files from %s.' % self.data_dir)
    self.features = discover_feature_shapes(self._cached_file_names[0])

  def _load_file(self, filename):
    """(x, x2, y, a) device tensors of one file, parsed once."""
    if filename in self._file_data:
      return self._file_data[filename]
    from telluride_decoding_b200 import tfrecord
    if set(self.in1_fields) - set(self.features.keys()):
      raise ValueError('Could not find all desired features (%s) in data (%s)' %
                       (self.in1_fields, self.features.keys()))
    groups = [list(self.in1_fields)]
    if self.out_field != 'ones':
      if self.out_field not in self.features:
        raise ValueError('Could not find %s in parsed_features[%s]' %
                         (self.out_field, self.features.keys()))
      groups.append([self.out_field])
    else:
      groups.append([])
    if self.in2_fields:
      for k in self.in2_fields:
        if k not in self.features:
          raise ValueError('Could not find %s in parsed_features[%s]' % (k, self.features.keys()))
    groups.append(list(self.in2_fields or []))
    if self.attended_field and self.attended_field not in self.features:
      raise ValueError('Could not find %s in parsed_features[%s]' %
                       (self.attended_field, self.features.keys()))
    groups.append([self.attended_field] if self.attended_field else [])
    names = [n for g in groups for n in g]
    widths = [int(self.features[n].shape[0]) for n in names]
    host = tfrecord.read_features(filename, names, widths)
    dev = _lib.as_device_f32(host.cuda(non_blocking=True))
    parts, col = [], 0
    for g in groups:
      w = sum(int(self.features[n].shape[0]) for n in g)
      parts.append(dev[:, col:col + w].contiguous() if w else None)
      col += w
    x, y, x2, a = parts
    n = int(x.shape[0])
    if y is None:                                   # 'ones' (brain_data.py:778-780)
      y = torch.ones((n, 1), dtype=torch.float32, device=x.device)
    if x2 is None:                                  # placeholder: first input column (:795-800)
      x2 = x[:, 0:1].contiguous()
    if a is None:                                   # placeholder zeros (:806-810)
      a = torch.zeros((n, 1), dtype=torch.float32, device=x.device)
    self._file_data[filename] = (x, x2, y, a)
    return self._file_data[filename]

  def add_trial(self, *args, **kwargs):
    raise TypeError('TFExampleData reads its trials from data_dir; use TrialBrainData.add_trial.')

  def set_trials(self, *args, **kwargs):
    raise TypeError('TFExampleData reads its trials from data_dir; use TrialBrainData.set_trials.')

  def create_dataset(self, mode='train', temporal_context=True, mixup_batch=False,
                     mismatch_batch=False):
    if mismatch_batch:
      raise NotImplementedError('mismatch_batch randomisation is outside the decoding hot path.')
    filename_list = self.filter_file_names(mode)
    if not filename_list:
      raise ValueError('No files to process in mode %s from directory %s: %s' %
                       (mode, self.data_dir, self.all_files()))
    loaded = [self._load_file(f) for f in filename_list]
    cat = lambda i: torch.cat([t[i] for t in loaded], dim=0)
    self._x, self._x2, self._y, self._a = cat(0), cat(1), cat(2), cat(3)
    files, cur = [], 0
    for f, t in zip(filename_list, loaded):
      files.append((f, cur, int(t[0].shape[0])))
      cur += int(t[0].shape[0])
    return self._make_dataset(mode, files, temporal_context, mixup_batch, offset_counts=False)


def create_brain_dataset(data_type, in_fields, out_field, frame_rate, pre_context=0,
                         post_context=0, in2_fields=None, in2_pre_context=0, in2_post_context=0,
                         input_offset=0, attended_field=None, initial_batch_size=1000000,
                         final_batch_size=1000, repeat_count=1, shuffle_buffer_size=1000,
                         data_dir=None, data_pattern='', train_file_pattern=None,
                         validate_file_pattern=None, test_file_pattern=None):
  """Creates any of the brain datasets that we know about (brain_data.py:959-1049).

  'test' -> TestBrainData, 'tfrecord' -> TFExampleData (files on disk, native parser),
  'trials' -> TrialBrainData (the same semantics for trials already in memory).
  """
  if not isinstance(data_type, str):
    raise TypeError('create_brain_dataset type must be a string.')
  if frame_rate <= 0:
    raise ValueError('frame_rate must be greater than 0.')
  kwargs = dict(pre_context=pre_context, post_context=post_context, in2_fields=in2_fields,
                in2_pre_context=in2_pre_context, in2_post_context=in2_post_context,
                input_offset=input_offset, attended_field=attended_field,
                initial_batch_size=initial_batch_size, final_batch_size=final_batch_size,
                repeat_count=repeat_count, shuffle_buffer_size=shuffle_buffer_size,
                data_dir=data_dir, data_pattern=data_pattern,
                train_file_pattern=train_file_pattern,
                validate_file_pattern=validate_file_pattern,
                test_file_pattern=test_file_pattern)
  if data_type == 'test':
    return TestBrainData(in_fields, out_field, frame_rate, **kwargs)
  if data_type in ('trials', 'memory'):
    return TrialBrainData(in_fields, out_field, frame_rate, **kwargs)
  if data_type in ('tfrecord', 'tfrecords', 'tfexample'):
    return TFExampleData(in_fields, out_field, frame_rate, **kwargs)
  raise TypeError('create_brain_dataset unknown data type %s' % data_type)
