"""Multi-GPU: time-sharded sufficient statistics + one allreduce.

The Gram is a sum over frames, so the frame axis is split into `world` contiguous, balanced
ranges (cut at 32-frame tile boundaries inside a trial; the lag halo around a cut is re-read from
the same trial, never exchanged and never zero padded).  Every rank runs the identical tile
schedule over its share and the partial statistics are combined with one NCCL allreduce of the
float32 Gram's upper triangle (plus one of the few float64 thin sums), issued through libtdgram.so's
own communicator (td_comm_*, include/tdgram.h).  One process per GPU; torch.distributed is the
bootstrap channel (and the whole transport in the gloo CPU tests).
"""

import ctypes

import torch
import torch.distributed as dist

from telluride_decoding_b200 import _lib


def shard_rows(seg_first, seg_keep, rank, world, align=32):
  """(first, keep) per segment for `rank`: a contiguous 1/world share of the rows
  [seg_first[s], seg_keep[s]) of all segments, balanced by row count.  Cuts that fall inside a
  segment are moved to a multiple of `align` rows so that both neighbours use whole k-tiles."""
  n = len(seg_keep)
  seg_first = [0] * n if seg_first is None else [int(v) for v in seg_first]
  seg_keep = [int(v) for v in seg_keep]
  sizes = [max(k - f, 0) for f, k in zip(seg_first, seg_keep)]
  total = sum(sizes)

  def cut(r):
    """Boundary r as (segment index, row within the segment)."""
    if r <= 0:
      return (0, seg_first[0] if n else 0)
    if r >= world:
      return (n, 0)
    target = total * r // world
    acc = 0
    for s in range(n):
      if acc + sizes[s] > target:
        row = seg_first[s] + (target - acc)
        row = (row + align // 2) // align * align          # nearest tile boundary
        row = min(max(row, seg_first[s]), seg_keep[s])
        return (s, row)
      acc += sizes[s]
    return (n, 0)

  (s0, r0), (s1, r1) = cut(rank), cut(rank + 1)
  first, keep = [], []
  for s in range(n):
    lo, hi = seg_first[s], seg_keep[s]
    if s < s0 or s > s1:
      lo = hi = seg_first[s]
    else:
      if s == s0:
        lo = max(lo, r0)
      if s == s1:
        hi = min(hi, r1)
    first.append(lo)
    keep.append(max(hi, lo))
  return first, keep


def is_distributed():
  return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class NativeComm(object):
  """NCCL communicator owned by libtdgram.so (td_comm_init / td_allreduce_* / td_comm_destroy,
  include/tdgram.h).  torch.distributed is only the bootstrap channel for the 128-byte unique id."""

  def __init__(self, group=None):
    _lib.require_cuda()
    self._lib = _lib.load()
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == 'nccl'
    ident = torch.zeros(128, dtype=torch.uint8)
    if self.rank == 0:
      buf = (ctypes.c_ubyte * 128)()
      _lib.check(self._lib.td_comm_unique_id(buf))
      ident = torch.tensor(list(buf), dtype=torch.uint8)
    if on_gpu:
      ident = ident.cuda()
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(ident, src=src, group=group)
    raw = (ctypes.c_ubyte * 128)(*[int(v) for v in ident.cpu().tolist()])
    self._comm = ctypes.c_void_p()
    _lib.check(self._lib.td_comm_init(ctypes.byref(self._comm), raw, self.rank, self.world))
    self._packed = None

  def allreduce_(self, t):
    """In-place sum of a contiguous float32 / float64 CUDA tensor."""
    if not (t.is_cuda and t.is_contiguous()):
      raise ValueError('NativeComm.allreduce_ needs a contiguous CUDA tensor')
    fn = {torch.float32: self._lib.td_allreduce_sum_f32,
          torch.float64: self._lib.td_allreduce_sum_f64}[t.dtype]
    with torch.cuda.device(t.device):
      _lib.check(fn(self._comm, _lib.ptr(t), ctypes.c_int64(t.numel()), _lib.current_stream_ptr()))
    return t

  def allreduce_sym_(self, g):
    """In-place sum of a symmetric float32 [n, n] CUDA matrix; only the upper triangle travels."""
    n = int(g.shape[0])
    if g.dim() != 2 or g.shape[1] != n or g.dtype != torch.float32 or g.stride(1) != 1:
      raise ValueError('NativeComm.allreduce_sym_ needs a square float32 matrix with unit column stride')
    need = n * (n + 1) // 2
    if self._packed is None or self._packed.numel() < need or self._packed.device != g.device:
      self._packed = None
      self._packed = torch.empty(need, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
      _lib.check(self._lib.td_allreduce_sym_f32(
          self._comm, _lib.ptr(g), ctypes.c_int64(n), ctypes.c_int64(g.stride(0)),
          _lib.ptr(self._packed), ctypes.c_size_t(self._packed.numel() * 4), _lib.current_stream_ptr()))
    return g

  def release_buffers(self):
    self._packed = None

  def close(self):
    if self._comm:
      _lib.check(self._lib.td_comm_destroy(self._comm))
      self._comm = ctypes.c_void_p()
    self._packed = None


_native_comms = {}


def native_comm(group=None):
  """The library-owned NCCL communicator of `group` (created on first use, collectively), or None
  when the job is not distributed over GPUs."""
  if not is_distributed() or not torch.cuda.is_available() or dist.get_backend(group) != 'nccl':
    return None
  key = id(group) if group is not None else 0
  if key not in _native_comms:
    _native_comms[key] = NativeComm(group)
  return _native_comms[key]


def close_native_comms():
  for comm in _native_comms.values():
    comm.close()
  _native_comms.clear()


def allreduce_sum_(tensors, group=None, symmetric=()):
  """In-place sum over ranks of a list of tensors.  float32 tensors (the Gram) are reduced as they
  are -- those also listed in `symmetric` (square, symmetric) send their upper triangle only --
  everything float64 is packed into ONE buffer first.  On GPUs the collectives run in
  libtdgram.so's own NCCL communicator; on CPU tensors (gloo tests) through torch.distributed."""
  if not is_distributed():
    return tensors
  comm = native_comm(group) if all(t.is_cuda for t in tensors) else None
  small = [t for t in tensors if t.dtype == torch.float64]
  for t in tensors:
    if t.dtype == torch.float64:
      continue
    if comm is not None and any(t is s for s in symmetric):
      comm.allreduce_sym_(t)
    elif comm is not None and t.is_contiguous():
      comm.allreduce_(t)
    else:
      dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
  if small:
    flat = torch.cat([t.reshape(-1) for t in small])
    if comm is not None:
      comm.allreduce_(flat)
    else:
      dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for t in small:
      t.copy_(flat[o:o + t.numel()].reshape(t.shape))
      o += t.numel()
  return tensors


def needed_row_range(seg_start, seg_len, seg_rows, pre, post, rank, world):
  """(row_begin, row_end) of a raw [frames, channels] array that `rank` touches: its share of the
  design rows (shard_rows over `seg_rows` kept rows per segment) plus the lag halo -- `pre` / `post`
  frames of the SAME segment on either side (brain_data.py:448-455 pads with zeros only at trial /
  chunk edges).  A rank whose share is empty gets (0, 0)."""
  first, keep = shard_rows(None, seg_rows, rank, world)
  lo, hi = None, 0
  for st, ln, f, k in zip(seg_start, seg_len, first, keep):
    if k <= f:
      continue
    a = int(st) + max(f - int(pre), 0)
    b = int(st) + min(k + int(post), int(ln))
    lo = a if lo is None else min(lo, a)
    hi = max(hi, b)
  return (0, 0) if lo is None else (lo, hi)


def _check_same_plan(dataset, group=None):
  """Raises unless every rank holds the same row plan (file order, kept rows per segment)."""
  import hashlib
  desc = repr(([f[0] for f in dataset.files], dataset.passes())).encode()
  mine = int(hashlib.sha1(desc).hexdigest()[:15], 16)
  t = torch.tensor([mine, -mine], dtype=torch.int64,
                   device='cuda' if dist.get_backend(group) == 'nccl' else 'cpu')
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  if int(t[0].item()) != -int(t[1].item()):
    raise ValueError('sharded statistics: the ranks disagree on the file order / kept rows of the '
                     'dataset; seed `random` identically on every rank before creating the BrainData '
                     'object (its file list is shuffled once, like the reference\'s).')


def sharded_regression_stats(dataset, gram_options=None, workspace=None, group=None, timings=None):
  """RegressionStats of the whole dataset, each rank accumulating its share of the frames.

  All ranks hold the same dataset description (the raw frames only need to be valid for the
  rank's own rows plus the lag halo, see needed_row_range).  `timings`, if a dict, receives CUDA
  events 'start', 'stats_done', 'allreduce_done' recorded on the current stream."""
  from telluride_decoding_b200 import brain_model
  from telluride_decoding_b200 import ops
  rank = dist.get_rank(group) if is_distributed() else 0
  world = dist.get_world_size(group) if is_distributed() else 1

  def mark(name):
    if timings is not None:
      ev = torch.cuda.Event(enable_timing=True)
      ev.record()
      timings[name] = ev

  x, y = dataset.stream('x'), dataset.stream('y')
  # Ranks share the frames in STORAGE order (segments sorted by their first raw row), which does not
  # depend on the once-shuffled file list (brain_data.py:206-207); which rows drop_remainder keeps
  # does depend on it, so the ranks must agree on that list (seed `random` identically).
  order = sorted(range(len(x.seg_start)), key=lambda i: x.seg_start[i])
  if world > 1:
    _check_same_plan(dataset, group)
  mark('start')
  gram = xtv = vtv = None
  for weight, keep in dataset.passes():
    f_sorted, k_sorted = shard_rows(None, [keep[i] for i in order], rank, world)
    first, mine = [0] * len(keep), [0] * len(keep)
    for pos, i in enumerate(order):
      first[i], mine[i] = f_sorted[pos], k_sorted[pos]
    g = ops.lagged_gram([x], mine, options=gram_options, workspace=workspace, seg_first=first)
    t, v = ops.lagged_thin(x, y, mine, seg_first=first)
    if gram is None:
      gram = g if weight == 1.0 else g * weight
      xtv, vtv = t * weight, v * weight
    else:
      gram = gram + g * weight
      xtv, vtv = xtv + t * weight, vtv + v * weight
  mark('stats_done')
  allreduce_sum_([gram, xtv, vtv], group=group, symmetric=(gram,))
  mark('allreduce_done')
  return brain_model.RegressionStats(gram, xtv, vtv)
