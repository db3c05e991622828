"""Streaming temporal context on the GPU against the unmodified reference Preprocessor.add_context
(tests/golden/streaming_context.npz, made by oracle/gen_golden.py) and the offline lag builder."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_numpy as ora
from telluride_decoding_b200 import preprocess

pytestmark = pytest.mark.gpu


def test_streaming_context_matches_reference(golden_dir):
  g = np.load(os.path.join(golden_dir, 'streaming_context.npz'))
  x = g['x']
  for tag in ('p2q3', 'p0q4', 'p3q0'):
    pre, post = [int(v) for v in g[tag + '_ctx']]
    cuts = [int(v) for v in g[tag + '_cuts']]
    p = preprocess.Preprocessor('test', 100, 100, pre_context=pre, post_context=post)
    outs = []
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
      got = p.add_context(x[a:b])
      assert got.is_cuda and got.dtype == torch.float32
      np.testing.assert_array_equal(got.cpu().numpy(), g['%s_%d' % (tag, i)].astype(np.float32))
      outs.append(got.cpu().numpy())
    np.testing.assert_array_equal(np.concatenate(outs), ora.add_context(x, pre, post)[:x.shape[0] - post])
    # reset starts a new stream
    p.context_reset()
    np.testing.assert_array_equal(p.process(x[:20]).cpu().numpy(), g[tag + '_0'].astype(np.float32)
                                  if cuts[1] == 20 else ora.add_context(x[:20], pre, post)[:20 - post])


def test_streaming_context_edge_cases():
  p = preprocess.Preprocessor('plain', 100, 100)
  x = torch.arange(12, dtype=torch.float32, device='cuda').reshape(6, 2)
  assert torch.equal(p.add_context(x), x)                       # no context: identity
  p = preprocess.Preprocessor('short', 100, 100, pre_context=1, post_context=4)
  with pytest.raises(ValueError, match='negative dimensions'):
    p.add_context(x[:2])                                        # like the reference's np.empty
  with pytest.raises(NotImplementedError):
    preprocess.Preprocessor('filt', 100, 100, highpass_cutoff=1.0)
  with pytest.raises(NotImplementedError):
    preprocess.Preprocessor('resample', 100, 50)
