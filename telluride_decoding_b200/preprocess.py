"""Streaming temporal context: the online counterpart of BrainData.add_temporal_context.

Mirrors `Preprocessor.add_context` / `context_reset` of telluride_decoding/preprocess.py:487-527:
batches of frames arrive one after the other, the last pre+post frames are carried as state, and
every call returns the lagged rows that have become complete -- row t of the running stream is
`[x[t-pre], ..., x[t], ..., x[t+post]]` (lag-major, channel-minor), delayed by `post_context`
frames; before the first frame the stream is zero (the initial state, :500-501).  The rows are built
on the GPU by td_lagged_rows_f32 from the carried state and the new batch; nothing else of the
reference's Preprocessor (filtering, resampling, normalisation) is on the decoding hot path.
"""
import torch

from telluride_decoding_b200 import _lib
from telluride_decoding_b200 import ops


class Preprocessor(object):
  """Context-only subset of preprocess.Preprocessor (same constructor keywords for the context)."""

  def __init__(self, name, fs_in=0.0, fs_out=0.0, pre_context=0, post_context=0, **kwargs):
    unsupported = {k: v for k, v in kwargs.items() if v}
    if unsupported or (fs_in and fs_out and fs_in != fs_out):
      raise NotImplementedError('Only the temporal context of Preprocessor is part of the B200 hot '
                                'path (got %s, fs_in=%g, fs_out=%g).' %
                                (sorted(unsupported.keys()), fs_in, fs_out))
    if pre_context < 0 or post_context < 0:
      raise ValueError('pre_context and post_context must be >= 0')
    self._name = name
    self._pre_context = int(pre_context)
    self._post_context = int(post_context)
    self._context_state = None

  @property
  def name(self):
    return self._name

  def context_reset(self):
    """Resets the saved state of the context (preprocess.py:524-527)."""
    self._context_state = None

  def add_context(self, data):
    """Lagged rows completed by this batch: float32 CUDA [rows, channels * (pre + 1 + post)].

    The first call returns `frames - post_context` rows, later calls `frames` rows (the last
    post_context samples appear in the next batch, preprocess.py:487-522)."""
    data = _lib.as_device_f32(data)
    if data.dim() != 2:
      raise ValueError('data must be 2-D (frames x channels), not %s' % (tuple(data.shape),))
    pre, post = self._pre_context, self._post_context
    if pre == 0 and post == 0:
      return data
    channels = int(data.shape[1])
    if self._context_state is None:
      self._context_state = torch.zeros((pre, channels), dtype=torch.float32, device=data.device)
    joined = torch.cat((self._context_state, data), dim=0).contiguous()
    keep = pre + post
    self._context_state = joined[joined.shape[0] - keep:].clone() if keep <= joined.shape[0] else joined
    rows = int(joined.shape[0]) - pre - post
    if rows < 0:       # the reference's np.empty raises the same way (preprocess.py:508)
      raise ValueError('negative dimensions are not allowed')
    if rows == 0:
      return torch.zeros((0, channels * (pre + 1 + post)), dtype=torch.float32, device=data.device)
    # Row t = [joined[t], ..., joined[t + pre + post]]: a post-only context over the joined piece.
    stream = ops.Stream(joined, 0, pre + post, [0], [int(joined.shape[0])])
    return ops.lagged_rows(stream, [rows], [0], rows)

  def process(self, data, reset=False):
    """Context only (preprocess.py:529-560 without the filters)."""
    if reset:
      self.context_reset()
    return self.add_context(data)
