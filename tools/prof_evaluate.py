"""Development: where does BrainModelLinearRegression.evaluate spend its time at the cfg 2 size?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from telluride_decoding_b200 import brain_data, brain_model, ops
x, y = bench.make_cfg2_host_data()
dev = torch.device('cuda', 0)
bd = brain_data.TrialBrainData('eeg', 'envelope', 128.0, pre_context=0, post_context=31, final_batch_size=768, shuffle_buffer_size=0)
bd.set_trials(['t%02d' % i for i in range(60)], [7680] * 60, torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev))
ds = bd.create_dataset('test')
model = brain_model.BrainModelLinearRegression(ds, regularization_lambda=0.1)
model.fit(ds)
def t(fn, n=5):
  fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): out = fn()
  torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, out
ms, pred = t(lambda: model._forward_dataset(ds)); print('forward (lagged_predict) %.3f ms' % ms)
ms, so = t(lambda: brain_model._stream_order(ds, pred)); print('stream order %.3f ms' % ms)
ms, truth = t(lambda: brain_model._stream_order(ds, ds.unlagged_rows('y')).contiguous()); print('truth rows %.3f ms' % ms)
ms, r = t(lambda: brain_model.batch_pearson(ds, truth, so)); print('batch pearson %.3f ms' % ms)
ms, _ = t(lambda: float(torch.mean(torch.square(truth - so)).item())); print('loss %.3f ms' % ms)
ms, res = t(lambda: model.evaluate(ds)); print('evaluate %.3f ms' % ms, res)
