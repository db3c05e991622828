"""BASELINE.json configs[3]: 306-ch MEG at 1 kHz x 100 lags (D = 30600) over 1 h, time-sharded Gram
+ NCCL allreduce.  Strong scaling: the same recording is split over the ranks.

  python tools/run_cfg4.py [--minutes 60] [--check]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_cfg4.py
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from telluride_decoding_b200 import ops, sharding

ap = argparse.ArgumentParser()
ap.add_argument('--minutes', type=float, default=60.0)
ap.add_argument('--channels', type=int, default=306)
ap.add_argument('--lags', type=int, default=100)
ap.add_argument('--check', action='store_true')
ap.add_argument('--reps', type=int, default=1)
args = ap.parse_args()
rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
  dist.init_process_group('nccl', device_id=dev)
trial_len = 100000
n_trials = max(1, int(round(args.minutes * 60 * 1000 / trial_len)))
c, L = args.channels, args.lags
gen = torch.Generator(device=dev); gen.manual_seed(4)
x = torch.randn((n_trials * trial_len, c), generator=gen, device=dev)   # same data on every rank
starts = [i * trial_len for i in range(n_trials)]
lens = [trial_len] * n_trials
stream = ops.Stream(x, 0, L - 1, starts, lens)
first, keep = sharding.shard_rows(None, lens, rank, world)
d = c * L
G = torch.empty((d, d), dtype=torch.float32, device=dev)
opt = ops.gram_options(timed=True)
times = []
for rep in range(args.reps + 1):
  if world > 1: dist.barrier()
  torch.cuda.synchronize(); t0 = time.perf_counter()
  ops.lagged_gram([stream], keep, out=G, options=opt, seg_first=first)
  torch.cuda.synchronize(); t1 = time.perf_counter()
  kernel_ms = ops.last_gram_kernel_ms()
  if world > 1:
    dist.all_reduce(G)
  torch.cuda.synchronize(); t2 = time.perf_counter()
  times.append((t1 - t0, t2 - t1, kernel_ms))
tg, tr, km = times[-1]
tt = torch.tensor([tg + tr, tg, tr, km * 1e-3], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
frames = n_trials * trial_len
res = dict(n_gpus=world, frames=frames, D=d, total_s=tt[0].item(), gram_call_s=tt[1].item(), allreduce_s=tt[2].item(),
           gram_kernel_s=tt[3].item(), frames_per_s=frames / tt[0].item(),
           tflops_alg=frames * d * (d + 1) / tt[0].item() / 1e12,
           tflops_alg_kernel_per_gpu=frames / world * d * (d + 1) / tt[3].item() / 1e12)
if args.check and rank == 0:
  # random lag blocks against float64 autocorrelations of the raw data
  x64 = x.view(n_trials, trial_len, c)
  errs = []
  for (k1, k2) in ((0, 0), (3, 50), (99, 99), (17, 18), (0, 99)):
    want = torch.zeros((c, c), dtype=torch.float64, device=dev)
    for i in range(n_trials):
      a = x64[i, k1:trial_len - (L - 1) + k1].to(torch.float64) if False else None
    # rows t in [0, n): X[t, k] = x[t + k] (zero beyond the trial end)
    for i in range(n_trials):
      n = trial_len
      hi = n - max(k1, k2)
      a = x64[i, k1:k1 + hi].to(torch.float64); b = x64[i, k2:k2 + hi].to(torch.float64)
      want += a.t() @ b
    got = G[k1 * c:(k1 + 1) * c, k2 * c:(k2 + 1) * c].to(torch.float64)
    errs.append(float(torch.linalg.norm(got - want) / torch.linalg.norm(want)))
  res['block_rel_err'] = errs
  res['symmetric'] = bool(torch.equal(G[:4096, :4096], G[:4096, :4096].t()))
if rank == 0:
  print(json.dumps(res))
if world > 1: dist.destroy_process_group()
