"""Development: random shapes through the tcgen05 Gram flavours (4 = operand built on chip, 5 = block-Toeplitz,
batched per-segment variants) against the fp32 CUDA-core cross-check kernel (method 1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from telluride_decoding_b200 import ops

def rel(a, b):
  return float(torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double()))

rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 120
bad = 0
t_start = time.time()
for case in range(n_cases):
  c = int(rng.choice([3, 8, 17, 24, 40, 64, 100, 306]))
  lags = int(rng.randint(2, 48)) if c < 200 else int(rng.randint(2, 9))
  pre = int(rng.randint(0, lags))
  post = lags - 1 - pre
  n_seg = int(rng.randint(1, 7))
  lens = [int(rng.randint(1, 4000)) for _ in range(n_seg)]
  keep = [int(rng.randint(max(1, n // 2), n + 1)) for n in lens]
  first = [int(rng.randint(0, max(1, k // 3))) for k in keep] if rng.rand() < 0.5 else None
  pad = int(rng.choice([0, 0, 1, 4]))
  x = torch.as_tensor((rng.randn(sum(lens), c + pad) + 0.1).astype(np.float32), device='cuda')[:, :c]
  starts = np.concatenate(([0], np.cumsum(lens)[:-1])).tolist()
  s = ops.Stream(x, pre, post, starts, lens)
  streams = [s]
  two = rng.rand() < 0.3
  if two:
    c2 = int(rng.choice([1, 5, 16]))
    l2 = int(rng.randint(1, 12))
    p2 = int(rng.randint(0, l2))
    x2 = torch.as_tensor(rng.randn(sum(lens), c2).astype(np.float32), device='cuda')
    streams.append(ops.Stream(x2, p2, l2 - 1 - p2, starts, lens))
  want = ops.lagged_gram(streams, keep, options=ops.gram_options(method=1), seg_first=first)
  for method in ((4, 6) if two else (4, 5, 6)):
    kw = {}
    if rng.rand() < 0.3:
      kw = dict(chunk_rows=int(rng.choice([32, 64, 128])), superchunk_rows=int(rng.choice([512, 1024, 4096])))
    got = ops.lagged_gram(streams, keep, options=ops.gram_options(method=method, cta_group=2, **kw), seg_first=first)
    e = rel(got, want) if float(want.abs().max()) > 0 else float(got.abs().max())
    if not (e < 1e-5) or not torch.equal(got, got.t()):
      bad += 1
      print('FAIL case %d method %d: c=%d lags=%d pre=%d lens=%s keep=%s first=%s two=%s kw=%s err=%.3e' %
            (case, method, c, lags, pre, lens, keep, first, two, kw, e), flush=True)
  if not two and n_seg > 1 and rng.rand() < 0.4:
    # batched per-segment Grams on equal segments
    n = int(rng.randint(40, 2000))
    xb = torch.as_tensor((rng.randn(n_seg * n, c) + 0.1).astype(np.float32), device='cuda')
    sb = ops.Stream(xb, pre, post, [i * n for i in range(n_seg)], [n] * n_seg)
    ref = torch.stack([ops.lagged_gram([ops.Stream(xb, pre, post, [i * n], [n])], [n], options=ops.gram_options(method=1)) for i in range(n_seg)])
    for method in (0, 4, 5):
      got = ops.lagged_gram_per_segment([sb], [n] * n_seg, options=ops.gram_options(method=method))
      e = rel(got, ref)
      if not (e < 1e-5):
        bad += 1
        print('FAIL batched case %d method %d: c=%d lags=%d pre=%d n=%d n_seg=%d err=%.3e' % (case, method, c, lags, pre, n, n_seg, e), flush=True)
torch.cuda.synchronize()
print('%d cases, %d failures, %.1f s' % (n_cases, bad, time.time() - t_start))
sys.exit(1 if bad else 0)
