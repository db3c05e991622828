"""BASELINE.json configs[0] shape: Telluride4-like ridge jackknife -- 32 trials of ~8000 frames at
100 Hz, 63 EEG channels -> 1 intensity, post_context = 20 (D = 1323), batch 100, lambda in 10^-6..10^0,
leave-one-trial-out: the reference fits 7 x 32 models, each with a full pass over the other 31 trials.
  python tools/bench_jackknife.py
Prints one JSON line: wall time of Regression.jackknife_over_regularizations and what it contains."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from telluride_decoding_b200 import brain_data, decoding, regression

rng = np.random.default_rng(1)
lens = [[8297, 8024, 8072, 7867][i % 4] for i in range(32)]
c, post = 63, 20
w_true = (rng.standard_normal((25, c)) / 10).astype(np.float32)
opts = decoding.DecodingOptions()
reg = regression.RegressionLinear(opts)
reg.preset_flags()
opts.post_context = post
bd = brain_data.TrialBrainData('eeg', 'intensity', 100.0, post_context=post, final_batch_size=opts.batch_size,
                               shuffle_buffer_size=opts.shuffle_buffer_size)
total = 0
for i, n in enumerate(lens):
  x = rng.standard_normal((n + 24, c)).astype(np.float32)
  y = sum(x[k:k + n] @ w_true[k] for k in range(25)).reshape(-1, 1) + 0.3 * rng.standard_normal((n, 1))
  bd.add_trial('trial_%02d' % i, x[:n], y.astype(np.float32))
  total += n
lambdas = regression.parse_regularization_values('normal')
for rep in range(2):
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  out = reg.jackknife_over_regularizations(opts, lambdas, brain_data_object=bd)
  torch.cuda.synchronize()
  wall = time.perf_counter() - t0
d = c * (post + 1)
best = max(out.items(), key=lambda kv: kv[1][0])
print(json.dumps(dict(workload='cfg1: 32 trials x ~8000 frames, 63 ch x 21 lags (D=%d), %d lambdas x 32 folds' % (d, len(lambdas)),
                      wall_s=wall, fits=len(lambdas) * 32, frames=total,
                      reference_equivalent_frames_per_s=total * 31 / 32.0 * len(lambdas) * 32 / wall,
                      best_lambda=float(best[0]), best_mean_r=float(best[1][0]))))
