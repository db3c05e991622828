"""Signal stores and window indexing (telluride_decoding/result_store.py).

`NumpyStore`, `WindowedDataStore` and `TwoResultStore` keep the reference's streaming interface
(add any number of frames, pull fixed-size windows) for callers that feed minibatches; the
index arithmetic they define -- windows [i*step, i*step + width) while at least `width` frames
remain (result_store.py:253-271), i.e. (n - width)//step + 1 windows -- is what the CUDA
windowed-correlation kernel reproduces in one pass (ops.windowed_corr).  Unlike the reference,
pulling a window does not memmove the whole buffer: a read cursor advances instead.
"""

import numpy as np


def window_starts(n_frames, window_width, window_step):
  """First frame of every window the reference would yield from n_frames frames."""
  if n_frames < window_width:
    return np.zeros((0,), dtype=np.int64)
  return np.arange((n_frames - window_width) // window_step + 1, dtype=np.int64) * window_step


class NumpyStore(object):
  """Grows as needed; `all_data` is everything added so far (result_store.py:36-163)."""

  def __init__(self, init_frame_count=10000, name='Generic'):
    if init_frame_count <= 0:
      raise ValueError('Initial frame count must be greater than 0, not %s' % (init_frame_count))
    self._init_frame_count = init_frame_count
    self._name = name
    self._data_store = None
    self._begin = 0      # read cursor (frames already consumed by next_window)
    self._end = 0        # write cursor

  @property
  def count(self):
    return self._end - self._begin

  @property
  def all_data(self):
    if self._data_store is None:
      return None
    return self._data_store[self._begin:self._end, :]

  def create_storage(self, data):
    if not isinstance(data, np.ndarray) or data.ndim != 2:
      raise TypeError('data must be a 2D numpy array, not %s' % type(data))
    if self._data_store is None:
      frames = max(self._init_frame_count, 2 * data.shape[0])
      self._data_store = np.zeros((frames, data.shape[1]))
      self._begin = self._end = 0
    elif self._data_store.shape[1] != data.shape[1]:
      raise ValueError('Data\'s shape has changed, and this is not allowed (%d to %d).' %
                       (self._data_store.shape[1], data.shape[1]))
    elif self._data_store.shape[0] < self._end + data.shape[0]:
      self._make_room(data.shape[0])

  def _make_room(self, incoming):
    live = self.count
    if self._begin >= incoming and self._begin * 2 >= self._data_store.shape[0]:
      # Enough consumed space at the front: compact instead of growing.
      self._data_store[:live, :] = self._data_store[self._begin:self._end, :]
    else:
      size = max(self._data_store.shape[0] * 2, live + 2 * incoming)
      grown = np.zeros((size, self._data_store.shape[1]))
      grown[:live, :] = self._data_store[self._begin:self._end, :]
      self._data_store = grown
    self._begin, self._end = 0, live

  def add_data(self, data):
    data = np.asarray(data)
    if data.ndim < 2:
      data = np.reshape(data, (-1, 1))
    self.create_storage(data)
    self._data_store[self._end:self._end + data.shape[0], :] = data
    self._end += data.shape[0]

  def next_window(self, window_size):
    """Yields the earliest `window_size` frames once (None if there are too few)."""
    if self.count < window_size:
      yield None
    else:
      p1 = np.copy(self._data_store[self._begin:self._begin + window_size, :])
      self._begin += window_size
      yield p1


class WindowedDataStore(NumpyStore):
  """Fixed-size windows every `window_step` frames (result_store.py:166-271)."""

  def __init__(self, window_step=100, window_width=None, pre_context=0, initial_frame_count=100):
    super(WindowedDataStore, self).__init__()
    if int(window_step) != window_step:
      raise ValueError('Must be an integer window_step for now, not %g.' % window_step)
    if window_width is None:
      window_width = int(3 * window_step)
    if window_step > window_width:
      raise ValueError('window_step (%d) must be less than or equal to '
                       'window_width (%d)' % (window_step, window_width))
    self._window_width = int(window_width)
    self._window_step = int(window_step)
    self._pre_context = int(pre_context)
    self._init_frame_count = int(initial_frame_count * max(window_step, window_width))

  def create_storage(self, data):
    first = self._data_store is None
    if not isinstance(data, np.ndarray) or len(data.shape) != 2:
      raise TypeError('data must be a 2D numpy array, not %s' % type(data))
    if not first and data.shape[1] != self._data_store.shape[1]:
      raise ValueError('Data\'s shape has changed, and this is not allowed.')
    super(WindowedDataStore, self).create_storage(data)
    if first and self._pre_context > 0:
      # Prime with zeros so windows can be centred on the sample time (result_store.py:241-244).
      if self._data_store.shape[0] < self._pre_context + data.shape[0]:
        self._data_store = np.zeros((self._pre_context + 2 * data.shape[0], data.shape[1]))
      self._end = self._pre_context

  def next_window(self):
    while self.count >= self._window_width:
      p1 = np.copy(self._data_store[self._begin:self._begin + self._window_width, :])
      self._begin += self._window_step
      yield p1


class TwoResultStore(object):
  """Two parallel WindowedDataStores (result_store.py:274-338)."""

  def __init__(self, window_width=100, window_step=100, pre_context=0, initial_frame_count=100):
    self._store1 = WindowedDataStore(window_step, window_width=window_width,
                                     pre_context=pre_context,
                                     initial_frame_count=initial_frame_count)
    self._store2 = WindowedDataStore(window_step, window_width=window_width, pre_context=0,
                                     initial_frame_count=initial_frame_count)

  @property
  def all_data(self):
    return self._store1.all_data, self._store2.all_data

  def add_data(self, s1, s2):
    if s1.shape[0] != s2.shape[0]:
      raise ValueError('Both data must have the same # frames, not %d vs. %d' %
                       (s1.shape[0], s2.shape[0]))
    self._store1.add_data(s1)
    self._store2.add_data(s2)

  def next_window(self):
    for p1 in self._store1.next_window():
      for p2 in self._store2.next_window():
        yield p1, p2
        break   # one window of the second signal per window of the first
