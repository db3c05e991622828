"""Development aid: wall-clock breakdown of one bench step (synchronising between phases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from telluride_decoding_b200 import brain_data, brain_model, ops

dev = torch.device('cuda', 0)
x_host, y_host = bench.make_host_data(1000)
x_pin, y_pin = torch.from_numpy(x_host).pin_memory(), torch.from_numpy(y_host).pin_memory()
fpt, trials = bench.CFG['frames_per_trial'], bench.CFG['trials']
method = int(sys.argv[1]) if len(sys.argv) > 1 else 0

def tic():
  torch.cuda.synchronize(); return time.perf_counter()

for it in range(4):
  t = [tic()]
  xd = x_pin.to(dev, non_blocking=True); yd = y_pin.to(dev, non_blocking=True); t.append(tic())
  bd = brain_data.TrialBrainData('eeg', 'envelope', 128.0, pre_context=0, post_context=31, final_batch_size=768, shuffle_buffer_size=0)
  for i in range(trials):
    bd.add_trial('trial_%02d' % i, xd[i * fpt:(i + 1) * fpt], yd[i * fpt:(i + 1) * fpt])
  ds = bd.create_dataset('train'); t.append(tic())
  x, y = ds.stream('x'), ds.stream('y')
  (w, keep), = ds.passes()
  g = ops.lagged_gram([x], keep); t.append(tic())
  tv, vv = ops.lagged_thin(x, y, keep); t.append(tic())
  stats = brain_model.RegressionStats(g, tv, vv)
  g_aug, c_aug = stats.augmented(True); n = stats.n; t.append(tic())
  sol = ops.ridge_solve(g_aug, c_aug, n, bench.LAMBDAS, method=method); t.append(tic())
  wh = sol.cpu(); t.append(tic())
  names = ['h2d', 'build', 'gram', 'thin', 'assemble', 'solve', 'd2h']
  print('iter %d: ' % it + '  '.join('%s %.2f ms' % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t[:-1], t[1:])) + '   total %.2f ms' % ((t[-1] - t[0]) * 1e3))
