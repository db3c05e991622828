"""Development: the halo kernel (method 4) against the fp32 cross-check kernel on a few shapes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from telluride_decoding_b200 import ops

def rel(a, b):
  return float(torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double()))

METHOD = int(sys.argv[1]) if len(sys.argv) > 1 else 4


def case(name, n_trials, frames, c, pre, post, keep_cut=0, first=None, c2=0, pre2=0, post2=0, ld_pad=0, **kw):
  if c2 and METHOD == 5:
    return True
  rng = np.random.RandomState(c * 7 + frames)
  x = (rng.randn(n_trials * frames, c + ld_pad) + 0.1).astype(np.float32)
  xd = torch.as_tensor(x, device='cuda')[:, :c]
  starts = [i * frames for i in range(n_trials)]
  lens = [frames] * n_trials
  keep = [frames - keep_cut * i for i in range(n_trials)]
  streams = [ops.Stream(xd, pre, post, starts, lens)]
  if c2:
    x2 = torch.as_tensor(rng.randn(n_trials * frames, c2).astype(np.float32), device='cuda')
    streams.append(ops.Stream(x2, pre2, post2, starts, lens))
  want = ops.lagged_gram(streams, keep, options=ops.gram_options(method=1), seg_first=first)
  t0 = time.time()
  got = ops.lagged_gram(streams, keep, options=ops.gram_options(method=METHOD, **kw), seg_first=first)
  torch.cuda.synchronize()
  e = rel(got, want)
  sym = bool(torch.equal(got, got.t()))
  print('%-28s D=%5d err %.2e sym %s  (%.2fs)' % (name, got.shape[0], e, sym, time.time() - t0), flush=True)
  return e < 1e-5 and sym

ok = True
ok &= case('c64_l32', 2, 3000, 64, 0, 31)
ok &= case('c64_l32_cut', 3, 3111, 64, 0, 31, keep_cut=7)
ok &= case('c40_l32_prepost', 3, 2500, 40, 5, 26, keep_cut=100)
ok &= case('c10_l60_unaligned_ld', 2, 2000, 10, 3, 56)
ok &= case('c306_l10', 2, 1500, 306, 0, 9)
ok &= case('c24_l23_first', 3, 4000, 24, 5, 17, first=[0, 1024, 37])
ok &= case('two_stream', 4, 2500, 40, 3, 22, keep_cut=33, c2=6, pre2=2, post2=5)
ok &= case('strided_view', 2, 2200, 60, 2, 9, ld_pad=4)
ok &= case('short_chunks', 5, 3111, 40, 0, 31, keep_cut=7, chunk_rows=32, superchunk_rows=1024)
ok &= case('lags_not_mult4', 2, 2048, 32, 1, 17)
print('ALL OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
