"""Randomised shapes through td_windowed_corr_f32 (every kernel path: vector rows of 4 / 8 / 16
columns, one contiguous column, generic; gcd(W, S) below and above a warp; strided inputs; several
segments and window shapes) against a plain numpy restatement of the reference's arithmetic
(infer_decoder.py:327-328, :441-450; result_store.py:262-271; brain_model.py:62-79)."""
import numpy as np
import pytest
import torch

from telluride_decoding_b200 import ops

pytestmark = pytest.mark.gpu


def _windows(n, w, s):
  return [] if n < w else [i * s for i in range((n - w) // s + 1)]


def _reference(a, b, starts, lens, widths, steps, mode, reduction, mu_a, mu_b, power):
  out = []
  for w, s in zip(widths, steps):
    for st, n in zip(starts, lens):
      aa, bb = a[st:st + n], (b[st:st + n] if b is not None else None)
      for i0 in _windows(n, w, s):
        xa = aa[i0:i0 + w]
        if mode == 'plain':
          out.append([np.mean(xa[:, 0].astype(np.float64))])
          continue
        xb = bb[i0:i0 + w]
        if mode == 'local':
          x64, y64 = xa.astype(np.float64), xb.astype(np.float64)
          xm, ym = x64 - x64.mean(0), y64 - y64.mean(0)
          xp, yp = (xm ** 2).sum(0), (ym ** 2).sum(0)
          if np.prod(xp) <= 0 or np.prod(yp) <= 0:
            out.append(np.zeros(xa.shape[1]))
          else:
            out.append((xm * ym).sum(0) / (np.sqrt(xp) * np.sqrt(yp)))
          continue
        c = ((xa - mu_a) * (xb - mu_b) / power).astype(np.float32)
        if mode == 'global_all':
          out.append(c.astype(np.float64).mean(0))
        elif reduction == 'first':
          out.append([c[:, 0].astype(np.float64).mean()])
        elif reduction == 'second':
          out.append([c[:, 1].astype(np.float64).mean()])
        elif reduction == 'mean':
          out.append([np.mean(c, axis=1).astype(np.float64).mean()])
        else:
          out.append([np.mean(np.sign(c) * c ** 2, axis=1).astype(np.float64).mean()])
  return np.asarray(out, dtype=np.float64)


@pytest.mark.parametrize('seed', range(6))
def test_random_window_shapes_and_layouts(seed):
  rng = np.random.RandomState(100 + seed)
  for _ in range(10):
    dims = int(rng.choice([1, 2, 3, 4, 5, 8, 16]))
    pad = int(rng.choice([0, 0, 0, 4, 3]))
    g = int(rng.choice([1, 4, 7, 32, 64, 128, 160]))
    n_seg = int(rng.randint(1, 5))
    lens = [int(rng.randint(1, 40)) * g + int(rng.randint(0, g)) for _ in range(n_seg)]
    total = sum(lens) + 8
    starts, cur = [], int(rng.choice([0, 3, 4]))
    for n in lens:
      starts.append(cur)
      cur += n
    widths = [g * int(rng.randint(1, 6)) for _ in range(int(rng.randint(1, 4)))]
    steps = [g * int(rng.randint(1, 4)) for _ in widths]
    a = (rng.randn(total, dims + pad) + 0.2).astype(np.float32)
    b = (0.5 * a + rng.randn(total, dims + pad)).astype(np.float32)
    mu_a = rng.randn(dims).astype(np.float32) * 0.1
    mu_b = rng.randn(dims).astype(np.float32) * 0.1
    power = (0.5 + rng.rand(dims)).astype(np.float32)
    ad = torch.as_tensor(a, device='cuda')[:, :dims]
    bd = torch.as_tensor(b, device='cuda')[:, :dims]
    dev = lambda v: torch.as_tensor(v, device='cuda')
    cases = [('plain', 'first'), ('local', 'first'), ('global_all', 'first'), ('global', 'first'), ('global', 'mean'),
             ('global', 'mean-squared')]
    if dims >= 2:
      cases.append(('global', 'second'))
    for mode, red in cases:
      got, off = ops.windowed_corr(ad, None if mode == 'plain' else bd, starts, lens, widths, steps, mode,
                                   reduction=red, mu_a=dev(mu_a), mu_b=dev(mu_b), power=dev(power))
      want = _reference(a[:, :dims], b[:, :dims], starts, lens, widths, steps, mode, red, mu_a, mu_b, power)
      assert got.shape[0] == want.shape[0] == off[-1], (mode, red, dims, g, lens, widths, steps)
      if want.shape[0]:
        tol = 1e-12 if red == 'first' and mode in ('plain', 'global', 'global_all') else 2e-6
        np.testing.assert_allclose(got.cpu().numpy(), want.reshape(got.shape), rtol=tol, atol=tol,
                                   err_msg=str((mode, red, dims, pad, g, lens, widths, steps)))
