"""Multi-GPU: time-sharded sufficient statistics + one allreduce.

The Gram is a sum over frames, so the frame axis is split into `world` contiguous, balanced
ranges (cut at 32-frame tile boundaries inside a trial; the lag halo around a cut is re-read from
the same trial, never exchanged and never zero padded).  Every rank runs the identical tile
schedule over its share and the partial statistics are combined with one NCCL allreduce of the
float32 Gram (plus one of the few float64 thin sums).  One process per GPU, torch.distributed
for the plumbing (nccl on GPUs, gloo in the CPU tests).
"""

import torch
import torch.distributed as dist


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


def allreduce_sum_(tensors, group=None):
  """In-place sum over ranks of a list of tensors: float32 tensors (the Gram) are reduced as they
  are, everything float64 is packed into ONE buffer first."""
  if not is_distributed():
    return tensors
  small = [t for t in tensors if t.dtype == torch.float64]
  for t in tensors:
    if t.dtype != torch.float64:
      dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
  if small:
    flat = torch.cat([t.reshape(-1) for t in small])
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
  allreduce_sum_([gram, xtv, vtv], group=group)
  mark('allreduce_done')
  return brain_model.RegressionStats(gram, xtv, vtv)
