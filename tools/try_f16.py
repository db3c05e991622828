"""Development: the fp16 x 3 Gram flavour (method 6) -- accuracy against float64 on a masked multi-trial
case and kernel time at the cfg 2 / cfg 4 widths, optionally from another build of the library
(TD_LIB_VARIANT=path/to/libtdgram_x.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from telluride_decoding_b200 import _lib
if os.environ.get('TD_LIB_VARIANT'):
  _lib.LIB_PATH = os.path.abspath(os.environ['TD_LIB_VARIANT'])
from telluride_decoding_b200 import ops

def lagged(x, pre, post):
  n, c = x.shape
  out = np.zeros((n, (pre + post + 1) * c))
  for k in range(pre + post + 1):
    sh = k - pre
    lo, hi = max(0, -sh), min(n, n - sh)
    out[lo:hi, k * c:(k + 1) * c] = x[lo + sh:hi + sh]
  return out

def check(lens, c, pre, post, methods=(4, 6), reps=3, chunk_rows=0):
  rng = np.random.RandomState(1)
  x = (rng.randn(sum(lens), c) + 0.1).astype(np.float32)
  xd = torch.as_tensor(x, device='cuda')
  starts = np.concatenate(([0], np.cumsum(lens)[:-1])).tolist()
  s = ops.Stream(xd, pre, post, starts, lens)
  design = np.concatenate([lagged(x[a:a + n].astype(np.float64), pre, post) for a, n in zip(starts, lens)])
  want = design.T @ design
  for m in methods:
    errs = []
    for _ in range(reps):
      got = ops.lagged_gram([s], lens, options=ops.gram_options(method=m, chunk_rows=chunk_rows)).cpu().numpy()
      errs.append(np.linalg.norm(got - want) / np.linalg.norm(want))
    print('[chunk_rows %d] ' % chunk_rows + 'lens %s c %d lags %d method %d: rel fro %s' % (lens, c, pre + post + 1, m, ' '.join('%.2e' % e for e in errs)), flush=True)

def timeit(name, n_trials, frames, c, lags, methods=(4, 6), reps=3, chunk_rows=0, superchunk_rows=0):
  gen = torch.Generator(device='cuda'); gen.manual_seed(1)
  x = torch.randn((n_trials * frames, c), generator=gen, device='cuda')
  starts = [i * frames for i in range(n_trials)]; lens = [frames] * n_trials
  s = ops.Stream(x, 0, lags - 1, starts, lens)
  d = c * lags
  G = torch.empty((d, d), device='cuda')
  f_alg = float(n_trials * frames) * d * (d + 1)
  ref = None
  for m in methods:
    opt = ops.gram_options(method=m, cta_group=2, timed=True, chunk_rows=chunk_rows, superchunk_rows=superchunk_rows)
    ts = []
    for _ in range(reps + 1):
      ops.lagged_gram([s], lens, out=G, options=opt)
      torch.cuda.synchronize()
      ts.append(ops.last_gram_kernel_ms())
    k = min(ts[1:])
    blk = G[:c, :c].double().clone()
    if ref is None: ref = blk
    print('[chunk_rows %d superchunk_rows %d] ' % (chunk_rows, superchunk_rows) + '%s method %d: kernel %.3f ms  %.1f TFLOP/s alg  (block diff vs first %.1e)' %
          (name, m, k, f_alg / k / 1e9, float((blk - ref).norm() / ref.norm())), flush=True)

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
print('library', _lib.LIB_PATH)
if what in ('all', 'check'):
  check([2000, 1777], 10, 3, 56)
  check([3000, 2500], 64, 0, 31)
if what == 'var':
  check([2000, 1777], 10, 3, 56, methods=(6,), reps=1)
  for cr in (256, 512):
    check([30000, 25000], 64, 0, 31, methods=(6,), reps=1, chunk_rows=cr)
  for cr in (128, 256):
    timeit('cfg2', 60, 7680, 64, 32, reps=4, chunk_rows=cr, methods=(6,))
    timeit('cfg4 slice', 3, 100000, 306, 100, reps=1, chunk_rows=cr, methods=(6,))
if what == 'power':
  import threading, time, pynvml
  pynvml.nvmlInit()
  h = pynvml.nvmlDeviceGetHandleByIndex(0)
  gen = torch.Generator(device='cuda'); gen.manual_seed(1)
  x = torch.randn((300000, 306), generator=gen, device='cuda')
  st = ops.Stream(x, 0, 99, [0, 100000, 200000], [100000] * 3)
  G = torch.empty((30600, 30600), device='cuda')
  for m, cr in ((4, 0), (6, 0), (6, 256)):
    opt = ops.gram_options(method=m, cta_group=2, timed=True, chunk_rows=cr)
    ops.lagged_gram([st], [100000] * 3, out=G, options=opt); torch.cuda.synchronize()
    samples = []; stop = [False]
    def sampler():
      while not stop[0]:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)
    th = threading.Thread(target=sampler); th.start()
    ts = []
    for _ in range(8):
      ops.lagged_gram([st], [100000] * 3, out=G, options=opt); torch.cuda.synchronize(); ts.append(ops.last_gram_kernel_ms())
    stop[0] = True; th.join()
    clk = sorted(c for c, _, _ in samples); pw = sorted(p for _, p, _ in samples)
    reasons = 0
    for _, _, r in samples: reasons |= r
    print('method %d chunk %d: kernel ms %s | SM clock median %d (min %d max %d) MHz, power median %.0f W (max %.0f), throttle bits 0x%x, %d samples' %
          (m, cr, ' '.join('%.0f' % t for t in ts), clk[len(clk) // 2], clk[0], clk[-1], pw[len(pw) // 2], pw[-1], reasons, len(samples)), flush=True)
if what == 'sc':
  for sc in (9984, 19968, 39936, 79872, 159744):
    timeit('cfg4 slice', 3, 100000, 306, 100, reps=1, methods=(6,), superchunk_rows=sc)
if what == 'quick':
  timeit('cfg2', 60, 7680, 64, 32, reps=5, methods=(6,))
  timeit('cfg4 slice', 3, 100000, 306, 100, reps=4, methods=(6,))
if what == 'prof':
  timeit('cfg4 slice', 3, 100000, 306, 100, reps=1, methods=(6,))
if what == 'prof2':
  timeit('cfg2', 60, 7680, 64, 32, reps=1, methods=(6,))
if what == 'chunks':
  for cr in (128, 256, 512, 1024):
    timeit('cfg2', 60, 7680, 64, 32, reps=4, chunk_rows=cr)
  for cr in (128, 256, 512):
    timeit('cfg4 slice', 3, 100000, 306, 100, reps=1, chunk_rows=cr, methods=(6,))
if what in ('all', 'time'):
  timeit('cfg2', 60, 7680, 64, 32, reps=4)
  timeit('cfg4 slice', 3, 100000, 306, 100, reps=2)
  timeit('cfg4 slice', 3, 100000, 306, 100, reps=2, chunk_rows=256, methods=(6,))
