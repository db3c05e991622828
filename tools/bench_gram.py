"""Development: Gram kernel time of the tcgen05 flavours (method 3 = staged planes + A in TMEM, 4 = operand
built on chip from raw tile + halo) on cfg2 and on a slice of cfg4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from telluride_decoding_b200 import ops

def run(name, n_trials, frames, c, lags, methods, reps=3, **kw):
  gen = torch.Generator(device='cuda'); gen.manual_seed(1)
  x = torch.randn((n_trials * frames, c), generator=gen, device='cuda')
  starts = [i * frames for i in range(n_trials)]; lens = [frames] * n_trials
  s = ops.Stream(x, 0, lags - 1, starts, lens)
  d = c * lags
  G = torch.empty((d, d), device='cuda')
  f_alg = float(n_trials * frames) * d * (d + 1)
  ref = None
  for m in methods:
    opt = ops.gram_options(method=m, cta_group=2, timed=True, **kw)
    ts = []
    for _ in range(reps + 1):
      t0 = time.perf_counter()
      ops.lagged_gram([s], lens, out=G, options=opt)
      torch.cuda.synchronize()
      ts.append((ops.last_gram_kernel_ms(), (time.perf_counter() - t0) * 1e3))
    k = min(t[0] for t in ts[1:]); w = min(t[1] for t in ts[1:])
    blk = G[:c, :c].clone()
    if ref is None: ref = blk
    print('%s method %d: kernel %.3f ms  call %.3f ms  %.1f TFLOP/s alg  (block diff vs first %.1e)' %
          (name, m, k, w, f_alg / k / 1e9, float((blk - ref).abs().max() / ref.abs().max())), flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
METHODS = [int(m) for m in sys.argv[2].split(',')] if len(sys.argv) > 2 else [3, 4]
if which in ('all', 'cfg2'):
  run('cfg2 64ch x 32 lags, 460800 frames', 60, 7680, 64, 32, METHODS, reps=5)
if which in ('all', 'cfg4'):
  run('cfg4 slice 306ch x 100 lags, 300000 frames', 3, 100000, 306, 100, METHODS, reps=2)
if which == 'sc':
  for sc in (0, 39200, 78400, 156800):
    print('superchunk_rows', sc)
    run('cfg4 slice 306ch x 100 lags, 300000 frames', 3, 100000, 306, 100, [4], reps=2, superchunk_rows=sc)
if which in ('all', 'cfg3'):
  run('cfg3-like 64ch x 40 lags', 60, 7680, 64, 40, [3, 4], reps=3)
