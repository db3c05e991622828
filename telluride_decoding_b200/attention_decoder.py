"""Attended-speaker decision rules (telluride_decoding/attention_decoder.py:116-173, :453-485).

Winner-take-all and the stepped (hysteresis) decoder.  The per-window comparison of the two
speakers' correlation means is `ops.wta_decide` on the GPU when whole experiments are decoded
at once; these classes keep the reference's streaming `attention(r1, r2)` interface.  The
state-space decoder (EM + Kalman, a serial scalar recursion) and the plotting helper are
outside the dense hot path.
"""

import numpy as np


class AttentionDecoder(object):
  """Winner takes all: is model 1 a better fit than model 2?  Strict greater-than."""

  def attention(self, r1, r2):
    return np.mean(r1) > np.mean(r2), 0, 0

  def tune(self, r1, r2):
    """An optional training step for tuning parameters."""
    del r1, r2


class StepAttentionDecoder(AttentionDecoder):
  """State starts at 0.5, moves +-0.1 per decision, clamped to [0.1, 0.9]; output state > 0.5."""

  def __init__(self):
    self.state = 0.5

  def attention(self, r1, r2):
    if np.mean(r1) > np.mean(r2):
      self.state = min(0.9, self.state + 0.1)
    else:
      self.state = max(0.1, self.state - 0.1)
    return self.state > 0.5, 0, 0


def step_decisions(wta):
  """The stepped decoder's outputs for a whole sequence of winner-take-all comparisons."""
  dec = StepAttentionDecoder()
  return np.array([dec.attention(1.0 if w else 0.0, 0.5)[0] for w in wta], dtype=bool)


def create_attention_decoder(type_name, window_step=100, frame_rate=100.0, ssd_offset=0.0):
  """'wta' or 'stepped' / 'step' (attention_decoder.py:453-485)."""
  del window_step, frame_rate, ssd_offset
  if type_name == 'wta':
    return AttentionDecoder()
  elif type_name == 'stepped' or type_name == 'step':
    return StepAttentionDecoder()
  elif type_name == 'ssd':
    raise ValueError('The state-space decoder (ssd) is outside the B200 decoding hot path.')
  raise ValueError('Unknown type (%s) requested from create_attention_decoder' % type_name)
