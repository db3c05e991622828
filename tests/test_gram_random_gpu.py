"""Randomised shapes through the tcgen05 Gram flavours (method 4, method 5, per-segment variants, two
streams, row ranges, unaligned inputs, tiny trials) against the fp32 CUDA-core cross-check kernel:
tools/stress_gram.py with a fixed seed."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_shapes_against_the_cross_check_kernel():
  res = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'stress_gram.py'), '7', '80'],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
  assert '0 failures' in res.stdout
