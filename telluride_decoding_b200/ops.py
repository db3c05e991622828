"""Python faces of the libtdgram.so entry points (one function per C symbol).

Tensors are torch CUDA tensors (they cross into the library as raw device pointers, i.e. what
DLPack hands over); segment tables, lambda lists and window shapes are host lists.  Each wrapper
cites the reference code the kernel replaces.
"""

import ctypes

import torch

from telluride_decoding_b200 import _lib
from telluride_decoding_b200._lib import Stream  # re-export  pylint: disable=unused-import

REDUCTIONS = {'first': 0, 'second': 1, 'mean': 2, 'mean-squared': 3, 'lda': 5}


def _streams_array(streams):
  arr = (_lib.TdStream * len(streams))()
  for i, s in enumerate(streams):
    arr[i] = s.c_struct()
  return arr


def gram_options(method=0, cta_group=0, chunk_rows=0, superchunk_rows=0, max_ctas=0, timed=False):
  opt = _lib.TdGramOptions(method, cta_group, chunk_rows, superchunk_rows, max_ctas)
  opt.reserved[0] = 1 if timed else 0
  return opt


def last_gram_kernel_ms():
  """Device time of the last tcgen05 Gram kernel launched with gram_options(timed=True)."""
  lib = _lib.load()
  lib.td_last_gram_kernel_ms.restype = ctypes.c_float
  return float(lib.td_last_gram_kernel_ms())


class GramWorkspace(object):
  """Caches the staging workspace between calls of the same shape (torch owns the memory)."""

  def __init__(self):
    self.buffer = None

  def get(self, nbytes, device):
    if self.buffer is None or self.buffer.numel() < nbytes or self.buffer.device != device:
      self.buffer = None
      self.buffer = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
    return self.buffer


_default_workspaces = {}


def _default_workspace_for(device):
  """The cached workspace of (device, current stream): calls on the same stream are ordered, calls
  on different streams must not share scratch memory.  ONE entry per device is kept (a call on
  another stream replaces it), so that dead or recycled stream handles cannot pile up buffers."""
  stream = torch.cuda.current_stream(device).cuda_stream
  entry = _default_workspaces.get(device.index)
  if entry is None or entry[0] != stream:
    entry = _default_workspaces[device.index] = (stream, GramWorkspace())
  return entry[1]


def release_workspaces():
  """Drops the cached Gram workspaces (tables; for method 3 also the staged planes of the recording)."""
  _default_workspaces.clear()


def lagged_gram(streams, seg_keep, out=None, options=None, workspace=None, seg_first=None):
  """G = Z'Z over the implicit lagged design Z = [X_0 | X_1]  (brain_model.py:437, cca.py:325-327).

  Design rows [seg_first, seg_keep) of each segment (seg_first None = 0).  Returns a float32
  [Dz, Dz] CUDA tensor (full symmetric)."""
  _lib.require_cuda()
  lib = _lib.load()
  device = streams[0].data.device
  dz = sum(s.width for s in streams)
  if out is None:
    out = torch.empty((dz, dz), dtype=torch.float32, device=device)
  options = options or gram_options()
  carr = _streams_array(streams)
  keep = _lib.i64_array(seg_keep)
  first = _lib.i64_array(seg_first) if seg_first is not None else None
  nbytes = ctypes.c_size_t(0)
  _lib.check(lib.td_lagged_gram_workspace_bytes(carr, len(streams), first, keep, len(seg_keep),
                                                ctypes.byref(options), ctypes.byref(nbytes)))
  ws = (workspace or _default_workspace_for(device)).get(nbytes.value, device)
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_gram_f32(carr, len(streams), first, keep, len(seg_keep), _lib.ptr(out),
                                      ctypes.c_int64(out.stride(0)), ctypes.byref(options),
                                      _lib.ptr(ws), ctypes.c_size_t(ws.numel()),
                                      _lib.current_stream_ptr()))
  return out


def lagged_gram_per_segment(streams, seg_keep, options=None, workspace=None, seg_first=None):
  """One Gram per segment, float32 [n_seg, Dz, Dz], from ONE launch of the tcgen05 kernel when every
  segment spans the same number of 32-frame k-tiles (leave-one-trial-out statistics: fold = total -
  trial, regression.py:197-214); otherwise one launch per segment."""
  _lib.require_cuda()
  lib = _lib.load()
  device = streams[0].data.device
  dz = sum(s.width for s in streams)
  n_seg = len(seg_keep)
  first_l = [0] * n_seg if seg_first is None else [int(v) for v in seg_first]
  tiles = set((int(k) + 31) // 32 - f // 32 if k > f else 0 for f, k in zip(first_l, seg_keep))
  out = torch.empty((n_seg, dz, dz), dtype=torch.float32, device=device)
  if len(tiles) != 1 or 0 in tiles:
    for i in range(n_seg):
      sub = [Stream(s.data, s.pre, s.post, [s.seg_start[i]], [s.seg_len[i]]) for s in streams]
      if seg_keep[i] > first_l[i]:
        lagged_gram(sub, [seg_keep[i]], out=out[i], options=options, workspace=workspace,
                    seg_first=[first_l[i]])
      else:
        out[i].zero_()
    return out
  options = options or gram_options()
  carr = _streams_array(streams)
  keep = _lib.i64_array(seg_keep)
  first = _lib.i64_array(seg_first) if seg_first is not None else None
  # method 5 (block-Toeplitz structure, one stream) or 4 (3xTF32) if asked for, else the default dense kernel (6: fp16 x 3)
  method = 5 if (options.method == 5 and len(streams) == 1) else (4 if options.method == 4 else 6)
  opt4 = _lib.TdGramOptions(method, 2, options.chunk_rows, 0, options.max_ctas)
  opt4.reserved[0] = options.reserved[0]
  nbytes = ctypes.c_size_t(0)
  _lib.check(lib.td_lagged_gram_batched_workspace_bytes(carr, len(streams), first, keep, n_seg,
                                                        ctypes.byref(opt4), ctypes.byref(nbytes)))
  ws = (workspace or _default_workspace_for(device)).get(nbytes.value, device)
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_gram_batched_f32(carr, len(streams), first, keep, n_seg, _lib.ptr(out),
                                              ctypes.c_int64(dz), ctypes.byref(opt4), _lib.ptr(ws),
                                              ctypes.c_size_t(ws.numel()), _lib.current_stream_ptr()))
  return out


def sum_batches(t):
  """float64 sum over the first axis of a contiguous float32 CUDA tensor [batches, ...] in one
  coalesced pass (the total Gram from the per-trial Grams)."""
  _lib.require_cuda()
  lib = _lib.load()
  t = t.contiguous()
  if t.dtype != torch.float32:
    return t.sum(dim=0, dtype=torch.float64)
  n = int(t[0].numel())
  out = torch.empty(t.shape[1:], dtype=torch.float64, device=t.device)
  with torch.cuda.device(t.device):
    _lib.check(lib.td_sum_batches_f32(_lib.ptr(t), ctypes.c_int64(n), ctypes.c_int64(n), int(t.shape[0]),
                                      _lib.ptr(out), _lib.current_stream_ptr()))
  return out


def lagged_gram_range(streams, seg_first, seg_keep, out=None):
  """Gram over design rows [seg_first, seg_keep) of each segment (CUDA cores; small corrections)."""
  _lib.require_cuda()
  lib = _lib.load()
  device = streams[0].data.device
  dz = sum(s.width for s in streams)
  if out is None:
    out = torch.empty((dz, dz), dtype=torch.float32, device=device)
  carr = _streams_array(streams)
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_gram_range_f32(carr, len(streams), _lib.i64_array(seg_first),
                                            _lib.i64_array(seg_keep), len(seg_keep), _lib.ptr(out),
                                            ctypes.c_int64(out.stride(0)),
                                            _lib.current_stream_ptr()))
  return out


def lagged_thin(x, v, seg_keep, seg_first=None):
  """(X'[v|1], [v|1]'[v|1]) in float64  (brain_model.py:436-439, cca.py:328-329)."""
  _lib.require_cuda()
  lib = _lib.load()
  device = x.data.device
  q = (v.channels if v is not None else 0) + 1
  xtv = torch.empty((x.width, q), dtype=torch.float64, device=device)
  vtv = torch.empty((q, q), dtype=torch.float64, device=device)
  xs = x.c_struct()
  vs = v.c_struct() if v is not None else None
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_thin_f64(ctypes.byref(xs), ctypes.byref(vs) if vs is not None else None,
                                      _lib.i64_array(seg_first) if seg_first is not None else None,
                                      _lib.i64_array(seg_keep), len(seg_keep), _lib.ptr(xtv),
                                      _lib.ptr(vtv), _lib.current_stream_ptr()))
  return xtv, vtv


def lagged_thin_per_segment(x, v, seg_keep, seg_first=None):
  """(X'[v|1] [n_seg, Dx, Q], [v|1]'[v|1] [n_seg, Q, Q]) in float64, one pair per segment."""
  _lib.require_cuda()
  lib = _lib.load()
  device = x.data.device
  q = (v.channels if v is not None else 0) + 1
  n_seg = len(seg_keep)
  xtv = torch.empty((n_seg, x.width, q), dtype=torch.float64, device=device)
  vtv = torch.empty((n_seg, q, q), dtype=torch.float64, device=device)
  xs = x.c_struct()
  vs = v.c_struct() if v is not None else None
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_thin_per_segment_f64(
        ctypes.byref(xs), ctypes.byref(vs) if vs is not None else None,
        _lib.i64_array(seg_first) if seg_first is not None else None, _lib.i64_array(seg_keep), n_seg,
        _lib.ptr(xtv), _lib.ptr(vtv), _lib.current_stream_ptr()))
  return xtv, vtv


def lagged_column_sums(x, seg_first, seg_keep, seg_slot, n_slots):
  """(sums float64 [n_slots, Dx], rows float64 [n_slots]): column sums of the implicit design over the
  row range of every segment, added into the segment's slot -- `np.sum(x, axis=0)` of every minibatch
  (brain_model.py:438) when the segments are the pieces of the minibatches."""
  _lib.require_cuda()
  lib = _lib.load()
  device = x.data.device
  sums = torch.empty((n_slots, x.width), dtype=torch.float64, device=device)
  rows = torch.empty((n_slots,), dtype=torch.float64, device=device)
  xs = x.c_struct()
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_column_sums_f64(ctypes.byref(xs), _lib.i64_array(seg_first), _lib.i64_array(seg_keep),
                                             _lib.i32_array(seg_slot), len(seg_keep), int(n_slots),
                                             _lib.ptr(sums), _lib.ptr(rows), _lib.current_stream_ptr()))
  return sums, rows


def lagged_lw_stat(x, seg_first, seg_keep, seg_slot, mean):
  """sum_t (sum_i (X[t, i] - mean[slot(t), i])^2)^2 as a float: the sum of all entries of the
  reference's `sum_x2tx2` (brain_model.py:440-443), which is all its Ledoit-Wolf estimate uses."""
  _lib.require_cuda()
  lib = _lib.load()
  device = x.data.device
  mean = mean.to(torch.float32).contiguous()
  out = torch.zeros((1,), dtype=torch.float64, device=device)
  xs = x.c_struct()
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_lw_stat_f64(ctypes.byref(xs), _lib.i64_array(seg_first), _lib.i64_array(seg_keep),
                                         _lib.i32_array(seg_slot), len(seg_keep), int(mean.shape[0]),
                                         _lib.ptr(mean), _lib.ptr(out), _lib.current_stream_ptr()))
  return float(out.item())


def lagged_predict_per_segment(x, seg_keep, out_row, weights, bias=None, out=None, total_rows=None):
  """out rows of segment s = X_s W[s] + b[s]: weights [n_seg, Dx, Dy], bias [n_seg, Dy] -- every
  held-out trial predicted by the model that never saw it."""
  _lib.require_cuda()
  lib = _lib.load()
  device = x.data.device
  weights = weights.to(torch.float32).contiguous()
  n_seg, dx, dy = (int(v) for v in weights.shape)
  if n_seg != len(seg_keep) or dx != x.width:
    raise ValueError('weights are %s, expected [%d, %d, Dy]' % (tuple(weights.shape), len(seg_keep), x.width))
  if bias is not None:
    bias = bias.to(torch.float32).contiguous()
  if out is None:
    if total_rows is None:
      total_rows = max([r + k for r, k in zip(out_row, seg_keep)] + [0])
    out = torch.empty((total_rows, dy), dtype=torch.float32, device=device)
  xs = x.c_struct()
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_predict_per_segment_f32(
        ctypes.byref(xs), _lib.i64_array(seg_keep), _lib.i64_array(out_row), n_seg, _lib.ptr(weights),
        _lib.ptr(bias), dy, _lib.ptr(out), ctypes.c_int64(out.stride(0)), _lib.current_stream_ptr()))
  return out


def lagged_predict(x, seg_keep, out_row, weights, bias=None, mean=None, out=None, total_rows=None):
  """out = (X - mean) W + b on the implicit design (brain_model.py:340-341, cca.py:157-158)."""
  _lib.require_cuda()
  lib = _lib.load()
  device = x.data.device
  weights = weights.contiguous()
  dy = int(weights.shape[1])
  if int(weights.shape[0]) != x.width:
    raise ValueError('weights have %d rows, the lagged input has %d columns' %
                     (weights.shape[0], x.width))
  if out is None:
    if total_rows is None:
      total_rows = max([r + k for r, k in zip(out_row, seg_keep)] + [0])
    out = torch.empty((total_rows, dy), dtype=torch.float32, device=device)
  xs = x.c_struct()
  with torch.cuda.device(device):
    _lib.check(lib.td_lagged_predict_f32(ctypes.byref(xs), _lib.i64_array(seg_keep),
                                         _lib.i64_array(out_row), len(seg_keep), _lib.ptr(weights),
                                         _lib.ptr(bias), _lib.ptr(mean), dy, _lib.ptr(out),
                                         ctypes.c_int64(out.stride(0)), _lib.current_stream_ptr()))
  return out


def lagged_rows(x, seg_keep, out_row, total_rows):
  """The lagged rows themselves, [total_rows, C*L]  (brain_data.py:453-455); tests / compat only."""
  _lib.require_cuda()
  lib = _lib.load()
  out = torch.zeros((total_rows, x.width), dtype=torch.float32, device=x.data.device)
  xs = x.c_struct()
  with torch.cuda.device(x.data.device):
    _lib.check(lib.td_lagged_rows_f32(ctypes.byref(xs), _lib.i64_array(seg_keep),
                                      _lib.i64_array(out_row), len(seg_keep), _lib.ptr(out),
                                      ctypes.c_int64(out.stride(0)), _lib.current_stream_ptr()))
  return out


def ridge_solve(g_aug, c_aug, n, lambdas, method=2):
  """[n_lambda, Da, Dy] float32 solutions of (G/n + lambda I) W = C/n  (brain_model.py:447-479)."""
  _lib.require_cuda()
  lib = _lib.load()
  g_aug = g_aug.to(torch.float64).contiguous()
  c_aug = c_aug.to(torch.float64).contiguous()
  da, dy = int(c_aug.shape[0]), int(c_aug.shape[1])
  lambdas = [float(l) for l in lambdas]
  out = torch.empty((len(lambdas), da, dy), dtype=torch.float32, device=g_aug.device)
  with torch.cuda.device(g_aug.device):
    _lib.check(lib.td_ridge_solve(_lib.ptr(g_aug), _lib.ptr(c_aug), da, dy, ctypes.c_double(n),
                                  _lib.f64_array(lambdas), len(lambdas), int(method),
                                  _lib.ptr(out), _lib.current_stream_ptr()))
  return out


def ridge_solve_batched(g_aug, c_aug, n, lambdas):
  """[n_batch, Da, Dy] float32 solutions of (G_b/n_b + lambda_b I) W_b = C_b/n_b for a batch of
  different systems in one batched Cholesky (the folds of a leave-one-trial-out fit)."""
  _lib.require_cuda()
  lib = _lib.load()
  g_aug = g_aug.to(torch.float64).contiguous()
  c_aug = c_aug.to(torch.float64).contiguous()
  nb, da, dy = int(c_aug.shape[0]), int(c_aug.shape[1]), int(c_aug.shape[2])
  n = [float(v) for v in n]
  lambdas = [float(l) for l in lambdas]
  if len(n) != nb or len(lambdas) != nb or tuple(g_aug.shape) != (nb, da, da):
    raise ValueError('ridge_solve_batched: inconsistent batch shapes')
  out = torch.empty((nb, da, dy), dtype=torch.float32, device=g_aug.device)
  with torch.cuda.device(g_aug.device):
    _lib.check(lib.td_ridge_solve_batched(_lib.ptr(g_aug), _lib.ptr(c_aug), da, dy, _lib.f64_array(n),
                                          _lib.f64_array(lambdas), nb, _lib.ptr(out),
                                          _lib.current_stream_ptr()))
  return out


def ridge_loo_solve(g_aug_total, c_aug_total, gram_fold, xtv_fold, vtv_fold, lamb, max_iter=40, tol=1e-10):
  """Every leave-one-trial-out fold from ONE Cholesky of the total system + a few fixed-point steps
  over the per-trial Grams (td_ridge_loo_solve).  Returns (W float32 [F, D+1, Dy], iterations,
  worst relative change of the last step); the caller falls back to ridge_solve_batched when that
  is not below its tolerance."""
  _lib.require_cuda()
  lib = _lib.load()
  g_aug_total = g_aug_total.to(torch.float64).contiguous()
  c_aug_total = c_aug_total.to(torch.float64).contiguous()
  gram_fold = gram_fold.to(torch.float32).contiguous()
  xtv_fold = xtv_fold.to(torch.float64).contiguous()
  vtv_fold = vtv_fold.to(torch.float64).contiguous()
  f, d = int(gram_fold.shape[0]), int(gram_fold.shape[1])
  dy = int(c_aug_total.shape[1])
  n_fold = [float(v) for v in vtv_fold[:, dy, dy].cpu().tolist()]
  n_total = float(g_aug_total[d, d].item())
  out = torch.empty((f, d + 1, dy), dtype=torch.float32, device=g_aug_total.device)
  iters = ctypes.c_int(0)
  worst = ctypes.c_double(0.0)
  with torch.cuda.device(g_aug_total.device):
    _lib.check(lib.td_ridge_loo_solve(_lib.ptr(g_aug_total), _lib.ptr(c_aug_total), _lib.ptr(gram_fold),
                                      _lib.ptr(xtv_fold), _lib.ptr(vtv_fold), _lib.f64_array(n_fold), d, dy, f,
                                      ctypes.c_double(n_total), ctypes.c_double(float(lamb)), int(max_iter),
                                      ctypes.c_double(float(tol)), _lib.ptr(out), ctypes.byref(iters),
                                      ctypes.byref(worst), _lib.current_stream_ptr()))
  return out, int(iters.value), float(worst.value)


def cca_solve(gxx, gyy, gxy, sum_x, sum_y, n_frames, n_norm, reg, eps_eig, dim):
  """(rot_x, rot_y, mean_x, mean_y, e) float32  (cca.py:337-367)."""
  _lib.require_cuda()
  lib = _lib.load()
  device = gxx.device
  f64 = lambda t: t.to(torch.float64).contiguous()
  gxx, gyy, gxy, sum_x, sum_y = f64(gxx), f64(gyy), f64(gxy), f64(sum_x).reshape(-1), f64(sum_y).reshape(-1)
  d1, d2 = int(gxx.shape[0]), int(gyy.shape[0])
  rot_x = torch.empty((d1, dim), dtype=torch.float32, device=device)
  rot_y = torch.empty((d2, dim), dtype=torch.float32, device=device)
  mean_x = torch.empty((1, d1), dtype=torch.float32, device=device)
  mean_y = torch.empty((1, d2), dtype=torch.float32, device=device)
  e = torch.empty((dim,), dtype=torch.float32, device=device)
  with torch.cuda.device(device):
    _lib.check(lib.td_cca_solve(_lib.ptr(gxx), _lib.ptr(gyy), _lib.ptr(gxy), _lib.ptr(sum_x),
                                _lib.ptr(sum_y), d1, d2, ctypes.c_double(n_frames),
                                ctypes.c_double(n_norm), ctypes.c_double(reg),
                                ctypes.c_double(eps_eig), int(dim), _lib.ptr(rot_x), _lib.ptr(rot_y),
                                _lib.ptr(mean_x), _lib.ptr(mean_y), _lib.ptr(e),
                                _lib.current_stream_ptr()))
  return rot_x, rot_y, mean_x, mean_y, e


def column_moments(a, b, sums=None):
  """sums[4, dims] (float64) += [sum a, sum b, sum a^2, sum b^2]  (infer_decoder.py:299-303)."""
  _lib.require_cuda()
  lib = _lib.load()
  dims = int(a.shape[1])
  if sums is None:
    sums = torch.zeros((4, dims), dtype=torch.float64, device=a.device)
  with torch.cuda.device(a.device):
    _lib.check(lib.td_column_moments_f64(_lib.ptr(a), ctypes.c_int64(a.stride(0)), _lib.ptr(b),
                                         ctypes.c_int64(b.stride(0)), ctypes.c_int64(a.shape[0]),
                                         dims, _lib.ptr(sums), _lib.current_stream_ptr()))
  return sums


def window_count(n_frames, width, step):
  return int(_lib.load().td_window_count(int(n_frames), int(width), int(step)))


def total_windows(seg_len, widths, steps):
  """Number of windows over all segments and shapes: sum of (n - W)//S + 1 where n >= W
  (result_store.py:262-271); vectorised, this sits on the host path of every decoder call."""
  import numpy as np
  lens = np.asarray(seg_len, dtype=np.int64)
  total = 0
  for w, s in zip(widths, steps):
    if w > 0 and s > 0:
      fits = lens[lens >= w]
      total += int(((fits - w) // s + 1).sum())
  return total


def windowed_corr(a, b, seg_start, seg_len, widths, steps, mode, reduction='first', mu_a=None,
                  mu_b=None, power=None, lda_params=None):
  """Sliding-window values for every (width, step) shape in one pass over the data.

  mode 'plain' | 'global' | 'local' | 'global_all' (see tdgram.h); reduction 'lda' takes
  lda_params = float64 CUDA [w_0..w_{dims-1}, slope, intercept].  Returns (values float64 [windows, cols],
  offsets list with n_sizes+1 entries); rows are ordered size, segment, window."""
  _lib.require_cuda()
  lib = _lib.load()
  mode_id = {'plain': 0, 'global': 1, 'local': 2, 'global_all': 3}[mode]
  dims = int(a.shape[1]) if mode_id else 1
  total = total_windows(seg_len, widths, steps)
  cols = dims if mode_id >= 2 else 1
  out = torch.empty((max(total, 1), cols), dtype=torch.float64, device=a.device)
  offsets = (ctypes.c_int64 * (len(widths) + 1))()
  with torch.cuda.device(a.device):
    _lib.check(lib.td_windowed_corr_f32(
        _lib.ptr(a), ctypes.c_int64(a.stride(0)), _lib.ptr(b),
        ctypes.c_int64(b.stride(0) if b is not None else 0), dims, _lib.i64_array(seg_start),
        _lib.i64_array(seg_len), len(seg_len), _lib.i64_array(widths), _lib.i64_array(steps),
        len(widths), mode_id, REDUCTIONS.get(reduction, 0), _lib.ptr(mu_a), _lib.ptr(mu_b),
        _lib.ptr(power), _lib.ptr(lda_params), _lib.ptr(out), offsets, _lib.current_stream_ptr()))
  return out[:total], [int(o) for o in offsets]


def wta_decide(r1, r2):
  """uint8 [n]: r1 > r2  (attention_decoder.py:134)."""
  _lib.require_cuda()
  lib = _lib.load()
  r1 = r1.to(torch.float64).contiguous().reshape(-1)
  r2 = r2.to(torch.float64).contiguous().reshape(-1)
  out = torch.empty((r1.numel(),), dtype=torch.uint8, device=r1.device)
  with torch.cuda.device(r1.device):
    _lib.check(lib.td_wta_decide_u8(_lib.ptr(r1), _lib.ptr(r2), ctypes.c_int64(r1.numel()),
                                    _lib.ptr(out), _lib.current_stream_ptr()))
  return out
