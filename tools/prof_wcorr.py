"""Development: times td_windowed_corr_f32 on the bandwidth-sized problem (bench leg `wcorr`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from telluride_decoding_b200 import ops
dims = int(sys.argv[1]) if len(sys.argv) > 1 else 8
red = sys.argv[2] if len(sys.argv) > 2 else 'mean'
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device('cuda', 0)
trials, frames = 1000, 76800
if dims == 1:
  frames *= 8
n = trials * frames
gen = torch.Generator(device=dev); gen.manual_seed(6)
a = torch.randn((n, dims), generator=gen, device=dev)
b = torch.randn((n, dims), generator=gen, device=dev)
mu = torch.zeros(dims, device=dev); pw = torch.ones(dims, device=dev) * 1.37
widths = [1280, 2560, 6400, 12800, 25600, 38400]
steps_w = [w // 2 for w in widths]
starts = [i * frames for i in range(trials)]
lens = [frames] * trials
def run():
  return ops.windowed_corr(a, b, starts, lens, widths, steps_w, 'global', reduction=red, mu_a=mu, mu_b=mu, power=pw)
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
  torch.cuda._sleep(3000000)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); v, _ = run(); e1.record(); torch.cuda.synchronize()
  ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
gb = (2.0 * n * dims * 4 + v.numel() * 8) / 1e9
print('dims %d red %s: %.3f ms  %.0f GB/s (%.2f of 6547)' % (dims, red, ms, gb / ms * 1e3, gb / ms * 1e3 / 6547))
