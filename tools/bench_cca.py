"""BASELINE.json configs[2]: CCA between 64-channel EEG (32 lags) and a 32-band spectrogram (16 lags),
1 h at 128 Hz, dim 8, reg 0.1: two-stream lagged Gram (D = 2048 + 512) + the CCA solve.
  python tools/bench_cca.py
Prints one JSON line with the device times of the statistics and of the solve."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from telluride_decoding_b200 import brain_data, cca

dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(3)
trials, fpt = 60, 7680
n = trials * fpt
eeg = torch.randn((n, 64), generator=gen, device=dev)
spec = torch.randn((n, 32), generator=gen, device=dev).abs()
spec = torch.nn.functional.avg_pool1d(spec.t().unsqueeze(0), 8, stride=1, padding=4)[0, :, :n].t().contiguous()
spec = (spec - spec.mean(0)) / spec.std(0)
spec[:, :4] += 0.3 * eeg[:, :4]                      # something to find
ones = torch.ones((n, 1), device=dev)
bd = brain_data.TrialBrainData('eeg', 'ones', 128.0, pre_context=0, post_context=31, in2_fields='spectrogram',
                               in2_pre_context=0, in2_post_context=15, final_batch_size=768,
                               shuffle_buffer_size=0)
bd.set_trials(['t%02d' % i for i in range(trials)], [fpt] * trials, eeg, ones, spec)
ds = bd.create_dataset('train')
times = []
for rep in range(4):
  e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
  e[0].record()
  stats = cca.CcaStats.from_dataset(ds)
  e[1].record()
  rot_x, rot_y, mean_x, mean_y, ev = cca.solve_cca(stats, 8, 0.1)
  e[2].record()
  torch.cuda.synchronize()
  times.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
stat_ms, solve_ms = times[-1]
d1, d2 = 2048, 512
flops = n * (d1 * (d1 + 1) + d2 * (d2 + 1) + 2 * d1 * d2)
print(json.dumps(dict(workload='cfg3: CCA 64ch x 32 lags vs 32 bands x 16 lags, 460800 frames, dim 8',
                      stats_ms=stat_ms, solve_ms=solve_ms, frames_per_s=n / ((stat_ms + solve_ms) * 1e-3),
                      tflops_alg_stats=flops / (stat_ms * 1e-3) / 1e12,
                      correlations=[float(v) for v in ev.reshape(-1)[:8]])))
