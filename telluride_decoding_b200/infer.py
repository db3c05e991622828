"""Attention-decoding test over window sizes (telluride_decoding/infer.py:327-464).

`run_reduction_test` takes the decoder object and the two speakers' test datasets instead of
directories of SavedModels / TFRecords, evaluates every window size in one pass, applies the
attention decoder and reports the fraction of correct windows exactly as infer.py:399-407.
"""

import numpy as np

from telluride_decoding_b200 import attention_decoder
from telluride_decoding_b200 import ops

DEFAULT_WINDOW_LIST = [10, 100, 200, 400, 700, 1000]      # infer.py:376


def find_first_segment(labels):
  """Frame at the end of the first run of equal labels (infer.py:301-324)."""
  if isinstance(labels, list):
    labels = np.asarray(labels)
  if not isinstance(labels, np.ndarray):
    raise TypeError('Labels input must be an ndarray, not %s' % type(labels))
  if labels.ndim != 1:
    raise TypeError('Labels input must be one-dimensional, not %s' % str(labels.shape))
  end_section = np.nonzero(np.logical_xor(labels, labels[0]))
  if end_section[0].shape[0]:
    return end_section[0][0]
  return 0


def calculate_time_axis(data, window_step, window_width, frame_rate):
  """Centre time (minutes) of each window (infer.py:172-199)."""
  if isinstance(data, (int, float)):
    num_points = int(data)
  elif isinstance(data, (list, np.ndarray)):
    num_points = len(data)
  else:
    raise TypeError('Unknown type passed as input argument.')
  return (np.arange(num_points) * window_step + window_width / 2.0) / frame_rate / 60.0


def run_reduction_test(model_object, bd1_test, bd2_test, decoder_type='wta',
                       window_list=None, frame_rate=100.0):
  """{window size: fraction of correctly decoded windows} (infer.py:327-464).

  bd1_test / bd2_test: test datasets whose output stream is speaker 1 / speaker 2's audio
  feature; the attention labels ride in the attended_speaker stream of bd2_test."""
  window_list = list(window_list or DEFAULT_WINDOW_LIST)
  res1 = model_object.windowed_results(bd1_test, window_list)
  res2 = model_object.windowed_results(bd2_test, window_list)
  out = {}
  for w in window_list:
    d1, _ = res1[w]
    d2, labels = res2[w]
    if decoder_type == 'wta' and len(d1):
      import torch
      att = ops.wta_decide(torch.as_tensor(d1, device='cuda'),
                           torch.as_tensor(d2, device='cuda')).cpu().numpy().astype(bool)
    else:
      decoder = attention_decoder.create_attention_decoder(decoder_type, window_step=w // 2,
                                                           frame_rate=frame_rate)
      end_first = find_first_segment(np.asarray(labels)) if len(labels) else 0
      if end_first:
        decoder.tune(d1[:end_first], d2[:end_first])
      att = np.array([decoder.attention(c1, c2)[0] for c1, c2 in zip(d1, d2)], dtype=bool)
    correct = np.logical_xor(att.reshape(-1, 1) >= 0.5, np.reshape(labels, (-1, 1)))
    out[w] = float(np.sum(correct)) / float(max(len(correct), 1))
  return out
