#!/usr/bin/env python
"""Benchmark of the dense decoding hot path (BASELINE.json metric: lagged-Gram frames/s + TFLOP/s).

Workload (BASELINE.json configs[1]): synthetic 64-channel EEG, 1 h at 128 Hz (60 trials x 7680
frames), 32 lags (pre 0 / post 31, D = 2048 + bias), one envelope target, ridge with a 20-lambda
regularisation sweep (logspace(-6, 3, 20)), minibatch 768 (no frame is dropped).

A step = one pass of the hot path: lagged Gram (tcgen05 3xTF32) + thin products X'[y|1] +
(N > 1: NCCL allreduce of the partial statistics) + the 20-lambda solve.  With N GPUs every rank
holds its own hour of recording (weak scaling: N hours fitted as one model; the lambda sweep is
split over ranks and all-gathered).

  python bench.py --gpus 1 --steps 5 --warmup 3            # this framework
  python bench.py --impl reference ...                     # the reference's CPU path (oracle port)
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

CFG = dict(workload='cfg2: 64-ch EEG, 1 h @128 Hz (60 trials x 7680), 32 lags, ridge, 20-lambda sweep',
           channels=64, trials=60, frames_per_trial=7680, pre=0, post=31, batch=768, n_lambda=20)
LAMBDAS = np.logspace(-6, 3, CFG['n_lambda'])
METRIC = 'lagged_gram_ridge_fit_frames_per_s'


def make_host_data(seed):
  """Synthetic z-scored EEG and an envelope that a 25-tap FIR of the EEG predicts (+ noise)."""
  rng = np.random.default_rng(seed)
  t = CFG['trials'] * CFG['frames_per_trial']
  x = rng.standard_normal((t, CFG['channels']), dtype=np.float32)
  taps = (rng.standard_normal((25, CFG['channels'])) / 40).astype(np.float32)
  y = np.zeros((t, 1), np.float32)
  for k in range(25):
    y[:t - k, 0] += x[k:] @ taps[k]
  y += 0.5 * rng.standard_normal((t, 1), dtype=np.float32)
  return x, y


# ----------------------------------------------------------------------------------------------
# Reference arm / CPU baseline: the reference's algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------

def cpu_reference(x, y, n_batches):
  """Times the reference's per-minibatch path (Hankel build + x'x + x'y, brain_data.py:446-455,
  brain_model.py:429-444) on `n_batches` minibatches and the full 20-lambda np.linalg.solve
  (brain_model.py:447-479); returns whole-workload frames/s with the Gram time scaled up."""
  from oracle import ref_numpy as ora
  b, pre, post = CFG['batch'], CFG['pre'], CFG['post']
  fpt = CFG['frames_per_trial']
  total_batches = CFG['trials'] * fpt // b
  n_batches = min(n_batches, total_batches)
  sum_xtx = sum_xty = 0.0
  sum_x = 0.0
  done = 0
  t0 = time.perf_counter()
  trial = 0
  while done < n_batches:
    xs, ys = x[trial * fpt:(trial + 1) * fpt], y[trial * fpt:(trial + 1) * fpt]
    design = ora.add_context(xs, pre, post)              # the reference materialises this
    for s in range(0, fpt - b + 1, b):
      if done >= n_batches:
        break
      xb = np.hstack((design[s:s + b], np.ones((b, 1), np.float32)))
      sum_xtx = sum_xtx + np.matmul(xb.T, xb)
      sum_x = sum_x + np.sum(xb, axis=0, keepdims=True)
      sum_xty = sum_xty + np.matmul(xb.T, ys[s:s + b])
      done += 1
    trial += 1
  t_gram = time.perf_counter() - t0
  n = done * b
  t0 = time.perf_counter()
  for lam in LAMBDAS:                                    # one solve per lambda, like the reference
    cov = sum_xtx / n
    cov += lam * np.identity(cov.shape[0])
    np.linalg.solve(cov, sum_xty / n)
  t_solve = time.perf_counter() - t0
  frames = CFG['trials'] * fpt
  # The reference re-accumulates the Gram for EVERY lambda (regression.py:381-409).
  t_one_pass = t_gram * total_batches / done
  t_step = CFG['n_lambda'] * t_one_pass + t_solve
  return dict(frames_per_s=frames / t_step, gram_pass_s=t_one_pass, solve_s=t_solve,
              batches_timed=done, total_batches=total_batches,
              gram_gflops=2.0 * n * 2049 ** 2 / t_gram / 1e9)


def cpu_cores():
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count()


def run_reference_arm(args, rank):
  if rank != 0:
    return
  x, y = make_host_data(1000)
  per_step = 12                                    # minibatches per step: a bounded sample
  for _ in range(args.warmup):
    cpu_reference(x, y, 2)
  t0 = time.perf_counter()
  res = None
  for _ in range(args.steps):
    res = cpu_reference(x, y, per_step)
  dt = (time.perf_counter() - t0) / max(args.steps, 1)
  sample = ('%d of %d minibatches of %d frames per step (Hankel build + x.T@x + x.T@y, numpy/BLAS '
            'float32) scaled to one pass, x %d passes (the reference recomputes the Gram per lambda, '
            'regression.py:381-409) + 20 x np.linalg.solve(2049); no tf.data overhead' %
            (res['batches_timed'], res['total_batches'], CFG['batch'], CFG['n_lambda']))
  line = dict(metric=METRIC, value=res['frames_per_s'], unit='frames/s', n_gpus=args.gpus,
              steps=args.steps, warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True,
              scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
              config=dict(CFG, parallelism='cpu'),
              cpu_baseline=dict(value=res['frames_per_s'], unit='frames/s', cores=cpu_cores(),
                                kind='port', sample=sample, gram_pass_s=res['gram_pass_s'],
                                solve_s=res['solve_s'], gram_gflops=res['gram_gflops']),
              e2e=dict(value=res['frames_per_s'], unit='frames/s', h2d_bytes_per_step=0,
                       d2h_bytes_per_step=0))
  print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# Clock sampling during the timed region
# ----------------------------------------------------------------------------------------------

class ClockSampler(object):
  QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.samples = []
    self.proc = None
    self.index = index

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
           '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
          stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) >= 7:
        self.samples.append((time.time(), parts))

  def stop(self, windows=None):
    """Median SM clock and throttle reasons of the samples taken inside `windows` (list of
    (t0, t1) wall-clock intervals = the timed regions); all samples if none fell inside."""
    if self.proc is None:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    inside = [p for t, p in self.samples
              if windows and any(a <= t <= b + 0.03 for a, b in windows)]
    in_window = bool(inside)
    for p in (inside or [p for _, p in self.samples]):
      try:
        sm.append(float(p[0])); mx.append(float(p[1]))
      except ValueError:
        continue
      for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                            'sw_power_cap'), p[3:7]):
        if val.lower().startswith('active'):
          reasons.add(name)
    return dict(sm_mhz=float(np.median(sm)) if sm else None,
                sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                samples=len(sm), sampled='inside the timed regions' if in_window else 'whole run')


# ----------------------------------------------------------------------------------------------
# This framework
# ----------------------------------------------------------------------------------------------

# dram__bytes_read.sum + dram__bytes_write.sum of one Gram launch on the bench workload (ncu --set
# full capture, profiles/r01_ncu_gram_kernel_summary.txt): 556.7 MB + 41.2 MB.  The staged planes are
# 4 x 118 MB and are read once (L2 hit rate 97 %), the Gram is written once.
GRAM_DRAM_TRAFFIC_BYTES = 597866752


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--solve-method', type=int, default=2,
                  help='2 batched blocked Cholesky sweep, 0 cuSOLVER Cholesky per lambda, 1 one eigh')
  ap.add_argument('--cta-group', type=int, default=0)
  ap.add_argument('--chunk-rows', type=int, default=0)
  ap.add_argument('--skip-cpu-baseline', action='store_true')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))

  if args.impl == 'reference':
    run_reference_arm(args, rank)
    return

  import torch
  import torch.distributed as dist
  from telluride_decoding_b200 import brain_data, brain_model, ops, sharding

  assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback)'
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  warmup = max(args.warmup, 3)

  x_host, y_host = make_host_data(1000 + rank)
  x_pin = torch.from_numpy(x_host).pin_memory()
  y_pin = torch.from_numpy(y_host).pin_memory()
  fpt, trials = CFG['frames_per_trial'], CFG['trials']
  frames = fpt * trials

  def build_data(xd, yd):
    bd = brain_data.TrialBrainData('eeg', 'envelope', 128.0, pre_context=CFG['pre'],
                                   post_context=CFG['post'], final_batch_size=CFG['batch'],
                                   shuffle_buffer_size=0)
    bd.set_trials(['trial_%02d' % i for i in range(trials)], [fpt] * trials, xd, yd)
    return bd

  gopt = ops.gram_options(cta_group=args.cta_group, chunk_rows=args.chunk_rows, timed=True)
  my_lambdas = [float(l) for i, l in enumerate(LAMBDAS) if i % world == rank]
  n_pad = (len(LAMBDAS) + world - 1) // world

  def fit(dataset):
    """The hot path: statistics -> (allreduce) -> lambda sweep.  Returns [n_lambda, D+1] weights."""
    stats = brain_model.RegressionStats.from_dataset(dataset, gram_options=gopt)
    if world > 1:
      sharding.allreduce_sum_([stats.gram, stats.xtv, stats.vtv])
    g_aug, c_aug = stats.augmented(True)
    sol = ops.ridge_solve(g_aug, c_aug, stats.n, my_lambdas or [1.0], method=args.solve_method)
    sol = sol[:, :, 0]
    if world > 1:
      pad = torch.zeros((n_pad, sol.shape[1]), dtype=sol.dtype, device=sol.device)
      pad[:len(my_lambdas)] = sol[:len(my_lambdas)]
      gathered = [torch.empty_like(pad) for _ in range(world)]
      dist.all_gather(gathered, pad)
      sol = torch.stack(gathered, dim=1).reshape(-1, sol.shape[1])[:len(LAMBDAS)]
    return sol

  # ---- resident arm: inputs already in HBM ------------------------------------------------------
  bd_res = build_data(x_pin.to(dev), y_pin.to(dev))
  ds_res = bd_res.create_dataset('train')
  assert ds_res.num_batches * CFG['batch'] == frames
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()                                # nvidia-smi needs ~0.2 s to deliver its first sample
  for _ in range(warmup):
    fit(ds_res)
  barrier()
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for _ in range(args.steps)]
  gram_ms = []
  barrier()
  wall0 = time.perf_counter()
  win0 = time.time()
  for i in range(args.steps):
    flush.zero_()                                  # evict L2 between timed iterations
    if world > 1:
      dist.barrier()
    ev[i][0].record()
    w = fit(ds_res)
    ev[i][1].record()
    gram_ms.append(ops.last_gram_kernel_ms())
  barrier()
  wall = time.perf_counter() - wall0
  windows = [(win0, time.time())]
  step_ms = [a.elapsed_time(b) for a, b in ev]
  t_res = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
  ms_per_step = float(t_res.item()) / args.steps

  # ---- end-to-end arm: host buffers in, weights out, copies inside the timed region -------------
  def e2e_step():
    # Host buffers straight into the public API: set_trials uploads the pinned arrays chunk by chunk
    # on a side stream and the statistics of the first trials start while the rest is in flight.
    ds = build_data(x_pin, y_pin).create_dataset('train')
    return fit(ds).cpu()

  for _ in range(2):
    e2e_step()
  barrier()
  e2e_ev = []
  win0 = time.time()
  for i in range(args.steps):
    flush.zero_()
    if world > 1:
      dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    w_host = e2e_step()
    b.record()
    e2e_ev.append((a, b))
  barrier()
  windows.append((win0, time.time()))
  clocks = sampler.stop(windows) if rank == 0 else None
  t_e2e = torch.tensor([sum(a.elapsed_time(b) for a, b in e2e_ev)], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
  e2e_ms = float(t_e2e.item()) / args.steps

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- roofline of the dominant kernel (the tcgen05 Gram) ---------------------------------------
  d = CFG['channels'] * (CFG['pre'] + 1 + CFG['post'])
  f_alg = float(frames) * d * (d + 1)               # symmetric-minimal flops of one launch
  gram_kernel_ms = float(np.mean(gram_ms))
  achieved = f_alg / (gram_kernel_ms * 1e-3) / 1e12
  peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(peaks_path):
    peaks = json.load(open(peaks_path))
    bf16 = peaks.get('bf16_tflops', 1590.0)
    peak_note = 'MEASURED_PEAKS.json bf16_tflops (burst) / 2 (tf32 rate) / 3 (3xTF32)'
  else:
    bf16 = 1590.0
    peak_note = 'fallback 1.59 PF bf16 / 2 / 3'
  peak = bf16 / 2.0 / 3.0
  # live TF32 dense rate for context (cuBLAS; denominator cross-check only)
  torch.backends.cuda.matmul.allow_tf32 = True
  a = torch.randn(8192, 8192, device=dev)
  b = torch.randn(8192, 8192, device=dev)
  best = 1e9
  for _ in range(6):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); torch.matmul(a, b); e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
  tf32_live = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
  torch.backends.cuda.matmul.allow_tf32 = False

  cpu = None
  if world == 1 and not args.skip_cpu_baseline:
    r = cpu_reference(x_host, y_host, 40)
    cpu = dict(value=r['frames_per_s'], unit='frames/s', cores=cpu_cores(), kind='port',
               sample=('%d of %d minibatches of %d frames (Hankel build + x.T@x + x.T@y) scaled to '
                       'one pass x %d lambda passes (regression.py:381-409) + 20 x np.linalg.solve; '
                       'numpy/BLAS float32, no tf.data overhead' %
                       (r['batches_timed'], r['total_batches'], CFG['batch'], CFG['n_lambda'])),
               gram_pass_s=r['gram_pass_s'], solve_s=r['solve_s'], gram_gflops=r['gram_gflops'],
               single_pass_frames_per_s=frames / (r['gram_pass_s'] + r['solve_s']))

  # stage, gram, symmetrize, thin, thin-reduce + solve: scale, colmajor, build, take_y, store and per
  # 64-column panel potf2 (factor) and one back-substitution kernel (cuBLAS GEMMs are not counted)
  n_panels = (CFG['channels'] * (CFG['pre'] + 1 + CFG['post']) + 1 + 63) // 64
  n_launch = 5 + 5 + 2 * n_panels

  # ---- AAD leg (BASELINE.json configs[4]): sliding-window correlation + winner-take-all ----------
  aad = None
  if world == 1:
    trials_a, frames_a = 1000, 7680                     # 60 s @128 Hz per trial
    na = trials_a * frames_a
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    env = torch.randn((na, 1), generator=gen, device=dev)
    p1 = (0.3 * env + torch.randn((na, 1), generator=gen, device=dev)).contiguous()   # attended
    p2 = torch.randn((na, 1), generator=gen, device=dev)                             # unattended
    mu = torch.zeros(1, device=dev); pw = torch.ones(1, device=dev)
    widths = [128, 256, 640, 1280, 2560, 3840]          # 1, 2, 5, 10, 20, 30 s windows
    steps_w = [w // 2 for w in widths]
    starts = [i * frames_a for i in range(trials_a)]
    lens = [frames_a] * trials_a
    def aad_step():
      r1, _ = ops.windowed_corr(env, p1, starts, lens, widths, steps_w, 'global', reduction='first',
                                mu_a=mu, mu_b=mu, power=pw)
      r2, _ = ops.windowed_corr(env, p2, starts, lens, widths, steps_w, 'global', reduction='first',
                                mu_a=mu, mu_b=mu, power=pw)
      return ops.wta_decide(r1, r2)
    for _ in range(3):
      dec = aad_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
      dec = aad_step()
    torch.cuda.synchronize()
    aad_ms = (time.perf_counter() - t0) / 5 * 1e3
    aad = dict(workload='cfg5: 1000 trials x 7680 frames, 2 speakers, windows 1/2/5/10/20/30 s (step W/2), '
                        'global-statistics correlation, reduction first, winner-take-all',
               windows_per_s=2 * int(dec.numel()) / (aad_ms * 1e-3), decisions=int(dec.numel()),
               ms=aad_ms, accuracy=float(dec.float().mean().item()),
               note='wall clock of the public ops incl. host-side table building; kernel-only numbers in profiles/')
  line = dict(
      metric=METRIC, value=world * frames / (ms_per_step * 1e-3), unit='frames/s', n_gpus=world,
      steps=args.steps, warmup=warmup, ms_per_step=ms_per_step, higher_is_better=True,
      scaling='weak', vs_baseline=None, dtype='tf32x3 (fp32-equivalent), f64 solve',
      data='synthetic',
      config=dict(CFG, parallelism='frames sharded x%d, 1 allreduce' % world if world > 1 else 'single GPU',
                  l2='256 MB write between timed steps (outside the per-step CUDA events)',
                  solve_method=args.solve_method, lambdas='logspace(-6,3,20)'),
      gpu_launches=n_launch * args.steps,
      tflops_alg=world * f_alg / (ms_per_step * 1e-3) / 1e12,
      gram_kernel_ms=gram_kernel_ms, step_wall_ms=wall / args.steps * 1e3,
      roofline=dict(bound='tensor', achieved=achieved, peak=peak, unit='TFLOP/s',
                    frac=achieved / peak, traffic=GRAM_DRAM_TRAFFIC_BYTES if world == 1 else None,
                    traffic_source='profiles/r01_ncu_gram_kernel_summary.txt (ncu --set full, same workload)',
                    kernel='ts::lagged_gram_ts_kernel',
                    flops_per_launch=f_alg, peak_source=peak_note, tf32_dense_live_tflops=tf32_live,
                    frac_of_live_tf32_over_3=achieved / (tf32_live / 3.0)),
      e2e=dict(value=world * frames / (e2e_ms * 1e-3), unit='frames/s',
               h2d_bytes_per_step=int(x_pin.numel() * 4 + y_pin.numel() * 4),
               d2h_bytes_per_step=int(w_host.numel() * 4), ms_per_step=e2e_ms),
      clocks=clocks, cpu_baseline=cpu, aad=aad)
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
