"""BASELINE.json configs[4]: two-speaker attention decoding, sliding-window correlation at 1-30 s
windows.  Measures the windowed-correlation kernel (memory bound) on synthetic model outputs:
  python tools/bench_aad.py [--dims 1] [--frames 7680] [--trials 1000]
Prints one JSON line: windows/s, algorithmic GB/s (8 B per frame per column read + 8 B per window
written) against the measured HBM copy bandwidth."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from telluride_decoding_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument('--dims', type=int, default=1)
ap.add_argument('--frames', type=int, default=7680)
ap.add_argument('--trials', type=int, default=1000)
ap.add_argument('--reduction', default='first')
ap.add_argument('--reps', type=int, default=20)
args = ap.parse_args()
dev = torch.device('cuda', 0)
n = args.trials * args.frames
gen = torch.Generator(device=dev); gen.manual_seed(5)
a = torch.randn((n, args.dims), generator=gen, device=dev)
b = (0.3 * a + torch.randn((n, args.dims), generator=gen, device=dev)).contiguous()
mu = torch.zeros(args.dims, device=dev); power = torch.ones(args.dims, device=dev)
widths = [128, 256, 640, 1280, 2560, 3840]          # 1, 2, 5, 10, 20, 30 s at 128 Hz
steps = [w // 2 for w in widths]
starts = [i * args.frames for i in range(args.trials)]
lens = [args.frames] * args.trials
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for rep in range(args.reps + 3):
  flush.zero_()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  vals, off = ops.windowed_corr(a, b, starts, lens, widths, steps, 'global', reduction=args.reduction,
                                mu_a=mu, mu_b=mu, power=power)
  e.record(); torch.cuda.synchronize()
  if rep >= 3: times.append(s.elapsed_time(e))
ms = float(np.median(times))
n_windows = int(vals.shape[0])
alg_bytes = 8.0 * n * args.dims + 8.0 * n_windows
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
hbm = json.load(open(peaks))['hbm_gbs'] if os.path.exists(peaks) else 6650.0
print(json.dumps(dict(kernel='windowed_corr_kernel', trials=args.trials, frames_per_trial=args.frames, dims=args.dims,
                      reduction=args.reduction, windows=n_windows, ms=ms, windows_per_s=n_windows / (ms * 1e-3),
                      frames_per_s=n / (ms * 1e-3), achieved_gbs=alg_bytes / (ms * 1e-3) / 1e9, peak_gbs=hbm,
                      frac=alg_bytes / (ms * 1e-3) / 1e9 / hbm, frac_of_8tbs=alg_bytes / (ms * 1e-3) / 1e9 / 8000.0)))
