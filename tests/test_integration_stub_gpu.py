"""The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would paste next to
brain_model.py:429-444) really runs against libtdgram.so and returns the reference's sums."""
import os
import re

import numpy as np
import pytest

from oracle import ref_numpy as ora
from tests import gpu_util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_stub_runs_and_matches_the_oracle():
  text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
  blocks = re.findall(r'```python\n(.*?)```', text, flags=re.S)
  stub = [b for b in blocks if 'def lagged_sufficient_statistics' in b]
  assert len(stub) == 1
  namespace = {}
  cwd = os.getcwd()
  os.chdir(ROOT)                       # the stub opens the library by its path relative to the repo root
  try:
    exec(compile(stub[0], 'INTEGRATION.md', 'exec'), namespace)   # pylint: disable=exec-used
    rng = np.random.RandomState(2)
    n, c, pre, post, kept = 3000, 8, 1, 6, 2900
    x = rng.randn(n, c).astype(np.float32)
    y = rng.randn(n, 1).astype(np.float32)
    gram, xtv, vtv = namespace['lagged_sufficient_statistics'](x, y, pre, post, kept)
  finally:
    os.chdir(cwd)
  design = ora.add_context(x, pre, post)[:kept].astype(np.float64)
  v = np.concatenate([y[:kept].astype(np.float64), np.ones((kept, 1))], axis=1)
  assert gpu_util.rel_fro(gram.cpu().numpy(), design.T @ design) < 1e-5
  np.testing.assert_allclose(xtv.cpu().numpy(), design.T @ v, rtol=1e-6, atol=1e-5)
  np.testing.assert_allclose(vtv.cpu().numpy(), v.T @ v, rtol=1e-9)
