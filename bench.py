#!/usr/bin/env python
"""Benchmark of the dense decoding hot path (BASELINE.json metric: lagged-Gram frames/s + TFLOP/s).

Headline workload = BASELINE.json configs[3], the configuration the north-star target is quoted on:
synthetic 306-channel MEG, 1 h at 1 kHz (36 trials x 100 000 frames), 100 lags (pre 0 / post 99,
D = 30 600), one envelope target, minibatch 1000 (no frame is dropped).

A step = one pass of the hot path over the WHOLE recording through the public API
(`sharding.sharded_regression_stats`): the lagged Gram X'X on tcgen05 tensor cores (split operands,
fp16 x 3 by default, 3xTF32 with --gram-method 4: fp32-equivalent products either way), the
thin products X'[y|1] in float64 and -- with N > 1 GPUs -- ONE NCCL allreduce of the partial
statistics.  Strong scaling: the SAME recording is cut into N contiguous frame ranges
(`sharding.shard_rows`, cuts on 32-frame tiles in the middle of trials, lag halo read from the same
trial), so `value` = frames of the one recording / step time.

Secondary legs (N = 1 only, after the headline): cfg2 (64-ch EEG x 32 lags, ridge with a 20-lambda
sweep, checked against the CPU oracle at every lambda), evaluate / predict throughput, cfg3 (CCA),
cfg5 (two-speaker attention decoding) and the windowed-correlation kernel's HBM roofline.

  python bench.py --gpus 1 --steps 3 --warmup 3            # this framework
  python bench.py --impl reference ...                     # the reference's CPU path (oracle port)
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
  # The CPU arm uses every host core it can; torchrun exports OMP_NUM_THREADS=1 to its workers, which
  # would silently halve (or worse) the baseline, so the BLAS thread count is pinned before numpy loads.
  try:
    _cores = len(os.sched_getaffinity(0))
  except AttributeError:
    _cores = os.cpu_count() or 1
  for _k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
    os.environ[_k] = str(_cores)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

CFG4 = dict(workload='cfg4: 306-ch MEG, 1 h @1 kHz (36 trials x 100000 frames), 100 lags (D = 30600): '
                     'lagged Gram + thin products X\'[y|1]; the whole recording every step',
            channels=306, trials=36, frames_per_trial=100000, pre=0, post=99, batch=1000)
CFG2 = dict(workload='cfg2: 64-ch EEG, 1 h @128 Hz (60 trials x 7680), 32 lags, ridge, 20-lambda sweep',
            channels=64, trials=60, frames_per_trial=7680, pre=0, post=31, batch=768, n_lambda=20)
LAMBDAS = np.logspace(-6, 3, CFG2['n_lambda'])
METRIC = 'lagged_gram_frames_per_s'
SEED = 1000


def cpu_cores():
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count()


def use_all_cores():
  """numpy/BLAS on every host core for the CPU legs (no-op if threadpoolctl is missing)."""
  try:
    import threadpoolctl
    threadpoolctl.threadpool_limits(limits=cpu_cores())
  except Exception:   # pylint: disable=broad-except
    pass


# ----------------------------------------------------------------------------------------------
# Synthetic recordings (seeded per trial, so that a rank can make just the trials it touches)
# ----------------------------------------------------------------------------------------------

def make_trial(cfg, i, n_taps=25, n_src=8):
  """z-scored sensor noise and an envelope that a short FIR of the first sensors predicts (+ noise)."""
  fpt, c = cfg['frames_per_trial'], cfg['channels']
  rng = np.random.default_rng(SEED + i)
  x = rng.standard_normal((fpt, c), dtype=np.float32)
  taps = (np.random.default_rng(SEED - 1).standard_normal((n_taps, n_src)) / 10).astype(np.float32)
  y = np.zeros((fpt, 1), np.float32)
  for k in range(n_taps):
    y[:fpt - k, 0] += x[k:, :n_src] @ taps[k]
  y += 0.5 * rng.standard_normal((fpt, 1), dtype=np.float32)
  return x, y


def fill_rows(cfg, x_out, y_out, row_lo, row_hi):
  """Rows [row_lo, row_hi) of the recording into x_out / y_out (which hold just those rows)."""
  fpt = cfg['frames_per_trial']
  for i in range(row_lo // fpt, (row_hi + fpt - 1) // fpt):
    x, y = make_trial(cfg, i)
    a, b = max(row_lo, i * fpt), min(row_hi, (i + 1) * fpt)
    x_out[a - row_lo:b - row_lo] = torch_or_numpy(x[a - i * fpt:b - i * fpt], x_out)
    y_out[a - row_lo:b - row_lo] = torch_or_numpy(y[a - i * fpt:b - i * fpt], y_out)


def torch_or_numpy(a, like):
  if isinstance(like, np.ndarray):
    return a
  import torch
  return torch.from_numpy(a)


def make_cfg2_host_data(seed=SEED):
  """Synthetic z-scored EEG and an envelope that a 25-tap FIR of the EEG predicts (+ noise)."""
  rng = np.random.default_rng(seed)
  t = CFG2['trials'] * CFG2['frames_per_trial']
  x = rng.standard_normal((t, CFG2['channels']), dtype=np.float32)
  taps = (rng.standard_normal((25, CFG2['channels'])) / 40).astype(np.float32)
  y = np.zeros((t, 1), np.float32)
  for k in range(25):
    y[:t - k, 0] += x[k:] @ taps[k]
  y += 0.5 * rng.standard_normal((t, 1), dtype=np.float32)
  return x, y


# ----------------------------------------------------------------------------------------------
# CPU legs: the reference's algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------

def cpu_cfg4_minibatches(n_batches, rows, keep_gram=False):
  """The reference's per-minibatch accumulation (lag design brain_data.py:446-455, then
  `sum_xtx += x.T @ x`, `sum_x`, `sum_xty`, brain_model.py:429-444, float32 numpy/BLAS) on the first
  `n_batches` minibatches of `rows` frames of trial 0 of the cfg4 recording.  Returns
  (seconds, frames, gram or None).  Nothing is extrapolated: frames / seconds is the rate."""
  from oracle import ref_numpy as ora
  pre, post = CFG4['pre'], CFG4['post']
  x, y = make_trial(CFG4, 0)
  need = n_batches * rows + post
  x, y = x[:need], y[:need]
  sum_xtx = sum_xty = sum_x = 0.0
  t0 = time.perf_counter()
  # One trial is one file: the context is added to the whole file and then batched; building it
  # per minibatch with the halo gives the same rows without holding 100000 x 30600 floats.
  for b in range(n_batches):
    lo, hi = b * rows, (b + 1) * rows
    src = x[max(lo - pre, 0):hi + post]
    design = ora.add_context(src, pre, post)[lo - max(lo - pre, 0):][:rows]
    xb = np.hstack((design, np.ones((rows, 1), np.float32)))
    sum_xtx = sum_xtx + np.matmul(xb.T, xb)
    sum_x = sum_x + np.sum(xb, axis=0, keepdims=True)
    sum_xty = sum_xty + np.matmul(xb.T, y[lo:hi])
  dt = time.perf_counter() - t0
  return dt, n_batches * rows, (sum_xtx if keep_gram else None)


def cpu_cfg2_full_pass(x, y, batch, max_batches=None):
  """One FULL pass of the reference's accumulation over the cfg2 recording (600 minibatches of 768,
  or 4608 of the regression preset's 100, regression.py:434) and the 20 solves
  (brain_model.py:447-479, float32 np.linalg.solve).  Returns times, weights [n_lambda, D+1]."""
  from oracle import ref_numpy as ora
  pre, post, fpt = CFG2['pre'], CFG2['post'], CFG2['frames_per_trial']
  sum_xtx = sum_xty = 0.0
  done = 0
  total = CFG2['trials'] * (fpt // batch)
  t0 = time.perf_counter()
  for trial in range(CFG2['trials']):
    xs, ys = x[trial * fpt:(trial + 1) * fpt], y[trial * fpt:(trial + 1) * fpt]
    design = ora.add_context(xs, pre, post)              # the reference materialises this
    for s in range(0, fpt - batch + 1, batch):
      if max_batches is not None and done >= max_batches:
        break
      xb = np.hstack((design[s:s + batch], np.ones((batch, 1), np.float32)))
      sum_xtx = sum_xtx + np.matmul(xb.T, xb)
      sum_xty = sum_xty + np.matmul(xb.T, ys[s:s + batch])
      done += 1
  t_gram = time.perf_counter() - t0
  n = done * batch
  weights = []
  t0 = time.perf_counter()
  for lam in LAMBDAS:                                    # one solve per lambda, like the reference
    cov = sum_xtx / n
    cov += (lam * np.identity(cov.shape[0])).astype(cov.dtype)
    weights.append(np.linalg.solve(cov, sum_xty / n)[:, 0])
  t_solve = time.perf_counter() - t0
  return dict(gram_s=t_gram, solve_s=t_solve, batches=done, total_batches=total, frames=n,
              weights=np.stack(weights), gram_gflops=2.0 * n * 2049 ** 2 / t_gram / 1e9)


def run_reference_arm(args, rank):
  """`--impl reference`: the reference's CPU implementation of the headline path on the host cores.
  One step = `--ref-batches` minibatches of `--ref-rows` frames of the cfg4 recording (a bounded
  sample: a whole pass is 3600 minibatches, hours of CPU); value = frames of the sample / its time."""
  if rank != 0:
    return
  use_all_cores()
  rows, nb = args.ref_rows, args.ref_batches
  for _ in range(args.warmup):
    cpu_cfg4_minibatches(1, min(rows, 200))
  t_sum = 0.0
  for _ in range(args.steps):
    dt, frames, _ = cpu_cfg4_minibatches(nb, rows)
    t_sum += dt
  dt = t_sum / max(args.steps, 1)
  value = nb * rows / dt
  d1 = CFG4['channels'] * (CFG4['pre'] + 1 + CFG4['post']) + 1
  sample = ('%d minibatch(es) of %d frames of trial 0 per step: lag design + x.T@x + sum(x) + x.T@y over '
            'D+1 = %d columns, numpy/BLAS float32 on %d threads; rate of the sample, nothing '
            'extrapolated; no tf.data overhead' % (nb, rows, d1, cpu_cores()))
  legs = {}
  if args.legs != 'none':
    x2, y2 = make_cfg2_host_data()
    r = cpu_cfg2_full_pass(x2, y2, CFG2['batch'])
    frames2 = CFG2['trials'] * CFG2['frames_per_trial']
    legs['cfg2'] = dict(
        workload=CFG2['workload'], measured='one full pass (600 minibatches of 768) + 20 solves',
        gram_pass_s=r['gram_s'], solve_s=r['solve_s'], gram_gflops=r['gram_gflops'],
        single_pass_frames_per_s=frames2 / (r['gram_s'] + r['solve_s']),
        x_n_lambda_frames_per_s=frames2 / (CFG2['n_lambda'] * r['gram_s'] + r['solve_s']),
        x_n_lambda_note='the reference re-accumulates the Gram for every lambda (regression.py:381-409): '
                        '20 x the measured pass + the 20 solves (derived, not re-run)')
    r100 = cpu_cfg2_full_pass(x2, y2, 100, max_batches=460)
    legs['cfg2_batch100'] = dict(
        measured='460 of 4560 minibatches of the regression preset batch 100 (regression.py:434)',
        extrapolated=True, gram_gflops=r100['gram_gflops'],
        single_pass_frames_per_s=frames2 / (r100['gram_s'] * r100['total_batches'] / r100['batches'] +
                                            r100['solve_s']))
  line = dict(metric=METRIC, value=value, unit='frames/s', n_gpus=args.gpus, steps=args.steps,
              warmup=args.warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling='strong',
              vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
              config=dict(CFG4, parallelism='cpu'),
              tflops_alg=value * (d1 - 1) * d1 / 1e12,
              cpu_baseline=dict(value=value, unit='frames/s', cores=cpu_cores(), kind='port', sample=sample,
                                gram_gflops=2.0 * value * d1 * d1 / 1e9),
              e2e=dict(value=value, unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
              legs=legs)
  print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# Clock sampling during the timed region
# ----------------------------------------------------------------------------------------------

class ClockSampler(object):
  QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index, period_ms=50):
    self.samples = []
    self.proc = None
    self.index = index
    self.period_ms = period_ms

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
           '--format=csv,noheader,nounits', '-lms', str(self.period_ms)], stdout=subprocess.PIPE,
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
    sm, mx, pw, reasons = [], [], [], set()
    inside = [p for t, p in self.samples
              if windows and any(a <= t <= b + 0.03 for a, b in windows)]
    in_window = bool(inside)
    for p in (inside or [p for _, p in self.samples]):
      try:
        sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
      except ValueError:
        continue
      for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                            'sw_power_cap'), p[3:7]):
        if val.lower().startswith('active'):
          reasons.add(name)
    return dict(sm_mhz=float(np.median(sm)) if sm else None,
                sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                power_w_median=float(np.median(pw)) if pw else None,
                samples=len(sm), sampled='inside the timed regions' if in_window else 'whole run')


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the Gram kernel on the full cfg 4 recording
# (ncu, mean of 5 launches of this very command, profiles/r02_ncu_launches_summary.txt): 1481 GB read +
# 308 GB written.  Algorithmic bytes are 8.15 GB (raw 4.41 + G 3.75): G (3.75 GB) fits no on-chip memory,
# so every (tile, superchunk) partial goes through L2 into HBM, and ~11 % of the raw boxes miss L2.
# 137 GB/s, 2 % of the HBM bandwidth of a tensor-bound kernel.
# fp16 x 3 (the default flavour): 1821 GB read + 329 GB written per launch, same reading (249 GB/s).
GRAM_DRAM_TRAFFIC_BYTES_CFG4 = {6: 2149602696704, 4: 1788808138547}      # by Gram method (6 = fp16 x 3, 4 = 3xTF32)


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    p = json.load(open(path))
    return dict(bf16=p.get('bf16_tflops', 1590.0), bf16_sustained=p.get('bf16_tflops_sustained', 1590.0),
                hbm=p.get('hbm_gbs', 6500.0), source='MEASURED_PEAKS.json')
  return dict(bf16=1590.0, bf16_sustained=1590.0, hbm=6500.0,
              source='fallback of /opt/skills/guides/B200_PROFILING.md (1.59 PF bf16, 6.5 TB/s)')


# ----------------------------------------------------------------------------------------------
# Secondary legs (N = 1)
# ----------------------------------------------------------------------------------------------

def timed_ms(fn, steps, flush=None, run_ahead=False):
  """Mean CUDA-event time of fn() over `steps` calls.  run_ahead: a ~1 ms spin kernel is queued first so
  that the host prepares and enqueues fn()'s work while the GPU is busy, and the events bracket the
  device work only (kernel duration, not launch latency)."""
  import torch
  ev = []
  for _ in range(steps):
    if flush is not None:
      flush.zero_()
    if run_ahead:
      torch.cuda._sleep(3000000)   # pylint: disable=protected-access
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    ev.append((a, b))
  torch.cuda.synchronize()
  return float(np.mean([a.elapsed_time(b) for a, b in ev])), out


def leg_cfg2(dev, peaks, flush, with_cpu):
  """BASELINE.json configs[1]: ridge fit with a 20-lambda sweep (Gram + thin sums + batched solve),
  resident and end to end; weights checked against the CPU oracle's full pass at EVERY lambda."""
  import torch
  from telluride_decoding_b200 import brain_data, brain_model, ops
  x_host, y_host = make_cfg2_host_data()
  x_pin = torch.from_numpy(x_host).pin_memory()
  y_pin = torch.from_numpy(y_host).pin_memory()
  fpt, trials = CFG2['frames_per_trial'], CFG2['trials']
  frames = fpt * trials

  def build_data(xd, yd, **kw):
    bd = brain_data.TrialBrainData('eeg', 'envelope', 128.0, pre_context=CFG2['pre'],
                                   post_context=CFG2['post'], final_batch_size=CFG2['batch'],
                                   shuffle_buffer_size=0, **kw)
    bd.set_trials(['trial_%02d' % i for i in range(trials)], [fpt] * trials, xd, yd)
    return bd

  gopt = ops.gram_options(timed=True)
  lambdas = [float(l) for l in LAMBDAS]

  def fit(dataset):
    stats = brain_model.RegressionStats.from_dataset(dataset, gram_options=gopt)
    g_aug, c_aug = stats.augmented(True)
    return ops.ridge_solve(g_aug, c_aug, stats.n, lambdas, method=2)[:, :, 0]

  ds_res = build_data(x_pin.to(dev), y_pin.to(dev)).create_dataset('train')
  for _ in range(3):
    fit(ds_res)
  gram_ms = []
  def fit_res():
    w = fit(ds_res)
    gram_ms.append(ops.last_gram_kernel_ms())
    return w
  step_ms, w = timed_ms(fit_res, 10, flush)
  def e2e_step():
    return fit(build_data(x_pin, y_pin).create_dataset('train')).cpu()
  for _ in range(2):
    e2e_step()
  e2e_ms, w_host = timed_ms(e2e_step, 10, flush)
  d = CFG2['channels'] * (CFG2['pre'] + 1 + CFG2['post'])
  f_alg = float(frames) * d * (d + 1)
  kernel_ms = float(np.mean(gram_ms))
  peak = peaks['bf16'] / 3.0                      # fp16 x 3: three kind::f16 MMAs per product
  leg = dict(workload=CFG2['workload'], ms_per_step=step_ms, frames_per_s=frames / (step_ms * 1e-3),
             e2e_ms_per_step=e2e_ms, e2e_frames_per_s=frames / (e2e_ms * 1e-3),
             h2d_bytes_per_step=int(x_pin.numel() * 4 + y_pin.numel() * 4),
             d2h_bytes_per_step=int(w_host.numel() * 4), gram_kernel_ms=kernel_ms,
             roofline=dict(bound='tensor', achieved=f_alg / (kernel_ms * 1e-3) / 1e12, peak=peak,
                           unit='TFLOP/s', frac=f_alg / (kernel_ms * 1e-3) / 1e12 / peak,
                           frac_of_3xtf32_peak=f_alg / (kernel_ms * 1e-3) / 1e12 / (peaks['bf16'] / 6.0),
                           peak_source=peaks['source'] + ' bf16_tflops (burst: a ms-long kernel; = the fp16 rate) / 3 '
                                       '(fp16 x 3 split); frac_of_3xtf32_peak is against bf16_tflops / 2 / 3'),
             lambdas='logspace(-6,3,20)', l2='256 MB write between timed steps')

  # evaluate / predict throughput: the only numbers the reference publishes
  # (doc/DecodingCodelab.md:654 ~100 k frames/s; TF2 demo notebook cell 20 ~32 k frames/s).
  bd_eval = build_data(x_pin.to(dev), y_pin.to(dev))
  ds_eval = bd_eval.create_dataset('test')
  model = brain_model.BrainModelLinearRegression(ds_eval, regularization_lambda=0.1)
  model.fit(ds_eval)
  for _ in range(2):
    model.evaluate(ds_eval)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(5):
    res = model.evaluate(ds_eval)
  torch.cuda.synchronize()
  ev_s = (time.perf_counter() - t0) / 5
  leg['evaluate'] = dict(frames_per_s=frames / ev_s, ms=ev_s * 1e3,
                         pearson_correlation_first=res['pearson_correlation_first'],
                         what='BrainModelLinearRegression.evaluate: lagged forward + per-minibatch Pearson r, '
                              'wall clock incl. the host read of the metrics',
                         reference_published='~100 k frames/s (doc/DecodingCodelab.md:654), ~32 k frames/s '
                                             '(TF2 demo notebook), other hardware')
  if with_cpu:
    use_all_cores()
    r = cpu_cfg2_full_pass(x_host, y_host, CFG2['batch'])
    w_gpu = w.cpu().numpy().astype(np.float64)
    w_cpu = r['weights'].astype(np.float64)
    w_rel = [float(np.linalg.norm(w_gpu[i] - w_cpu[i]) / np.linalg.norm(w_cpu[i]))
             for i in range(len(lambdas))]
    leg['parity'] = dict(w_rel_max=max(w_rel), w_rel_per_lambda=w_rel, tolerance=1e-4,
                         against='oracle port, full float32 pass + np.linalg.solve at every lambda')
    leg['cpu_baseline'] = dict(
        kind='port', cores=cpu_cores(), measured='one full pass (600 minibatches of 768) + 20 solves',
        gram_pass_s=r['gram_s'], solve_s=r['solve_s'], gram_gflops=r['gram_gflops'],
        single_pass_frames_per_s=frames / (r['gram_s'] + r['solve_s']),
        x_n_lambda_frames_per_s=frames / (CFG2['n_lambda'] * r['gram_s'] + r['solve_s']),
        x_n_lambda_note='derived: 20 x the measured pass (regression.py:381-409) + the solves')
  return leg


def leg_windowed_corr(dev, peaks):
  """The sliding-window correlation kernel alone on a bandwidth-sized problem (1000 trials x 76 800
  frames x 8 correlation columns, six window shapes, global statistics): HBM GB/s against the
  measured copy bandwidth and against the 8 TB/s north_star quotes."""
  import torch
  from telluride_decoding_b200 import ops
  trials, frames, dims = 1000, 76800, 8
  n = trials * frames
  gen = torch.Generator(device=dev); gen.manual_seed(6)
  a = torch.randn((n, dims), generator=gen, device=dev)
  b = torch.randn((n, dims), generator=gen, device=dev)
  mu = torch.zeros(dims, device=dev); pw = torch.ones(dims, device=dev)
  widths = [1280, 2560, 6400, 12800, 25600, 38400]
  steps_w = [w // 2 for w in widths]
  starts = [i * frames for i in range(trials)]
  lens = [frames] * trials
  def run():
    return ops.windowed_corr(a, b, starts, lens, widths, steps_w, 'global', reduction='mean',
                             mu_a=mu, mu_b=mu, power=pw)
  for _ in range(3):
    run()
  ms, (vals, _) = timed_ms(run, 10, run_ahead=True)
  bytes_alg = 2.0 * n * dims * 4 + vals.numel() * 8
  gbs = bytes_alg / (ms * 1e-3) / 1e9
  del a, b
  return dict(workload='1000 trials x 76800 frames x 8 columns, 6 window shapes (10-300 s @128 Hz), '
                       'reduction mean; inputs 4.9 GB >> L2',
              ms=ms, bytes_per_launch=bytes_alg,
              roofline=dict(bound='hbm', achieved=gbs, peak=peaks['hbm'], unit='GB/s',
                            frac=gbs / peaks['hbm'], frac_of_8tbs=gbs / 8000.0,
                            peak_source=peaks['source'] + ' hbm_gbs',
                            timing='CUDA events around the device work of the public op (table uploads + partial-sum kernel + window kernel), enqueued behind a spin kernel so that host latency is excluded'))


def leg_cfg3(dev):
  """BASELINE.json configs[2]: BrainModelCCA, 64-ch EEG x 32 lags vs 32-band spectrogram x 16 lags,
  8 canonical components, 1 h at 128 Hz."""
  import torch
  from telluride_decoding_b200 import brain_data, cca
  rng = np.random.default_rng(SEED + 3)
  trials, fpt = 60, 7680
  n = trials * fpt
  x = rng.standard_normal((n, 64), dtype=np.float32)
  mix = (rng.standard_normal((64, 32)) / 8).astype(np.float32)
  x2 = (x @ mix + rng.standard_normal((n, 32), dtype=np.float32)).astype(np.float32)
  y = np.zeros((n, 1), np.float32)
  bd = brain_data.TrialBrainData('eeg', 'envelope', 128.0, pre_context=0, post_context=31,
                                 in2_fields='spectrogram', in2_pre_context=0, in2_post_context=15,
                                 final_batch_size=768, shuffle_buffer_size=0)
  bd.set_trials(['t%02d' % i for i in range(trials)], [fpt] * trials, torch.from_numpy(x).to(dev),
                torch.from_numpy(y).to(dev), input2_data=torch.from_numpy(x2).to(dev))
  ds = bd.create_dataset('train')
  model = cca.BrainModelCCA(ds, cca_dims=8, regularization_lambda=0.1)
  for _ in range(2):
    model.fit(ds)
  ms, _ = timed_ms(lambda: model.fit(ds), 5)
  def stats_only():
    return cca.CcaStats.from_dataset(ds)
  ms_stats, _ = timed_ms(stats_only, 5)
  res = model.evaluate(ds)
  return dict(workload='cfg3: CCA 64 ch x 32 lags (2048) vs 32 bands x 16 lags (512), dim 8, 460800 frames',
              fit_ms=ms, stats_ms=ms_stats, solve_ms=ms - ms_stats, frames_per_s=n / (ms * 1e-3),
              first_canonical_r=float(res['cca_pearson_correlation_first']))


def leg_cfg5(dev, n_trials=1000):
  """BASELINE.json configs[4] for real: two-speaker attention decoding, 1000 trials x 60 s @128 Hz,
  64 channels x 32 lags, a leave-one-trial-out ridge decoder (1000 folds), sliding-window correlation
  at 1-30 s windows, winner-take-all -- infer.leave_one_trial_out_test, one sweep over the data."""
  import torch
  from telluride_decoding_b200 import brain_data, infer
  frames, c, lags = 7680, 64, 32
  n = n_trials * frames
  gen = torch.Generator(device=dev); gen.manual_seed(5)
  s1 = torch.randn((n, 1), generator=gen, device=dev)
  s2 = torch.randn((n, 1), generator=gen, device=dev)
  label = (torch.arange(n, device=dev) // frames % 2).to(torch.float32).reshape(-1, 1)   # alternates per trial
  att = torch.where(label == 0, s1, s2)
  un = torch.where(label == 0, s2, s1)
  eeg = torch.randn((n, c), generator=gen, device=dev)
  taps = torch.randn((8, c), generator=gen, device=dev) / 100  # weak drive (r ~ 0.2, like EEG): 1 s windows are hard, 30 s easy
  drive = att + 0.1 * un                                   # the unattended speaker at 0.1 gain
  for k in range(8):                                       # the EEG lags the audio by 0..7 frames
    eeg[k:] += drive[:n - k] * taps[k]
  del un, drive
  bd = brain_data.TrialBrainData('eeg', 'attended', 128.0, pre_context=0, post_context=lags - 1,
                                 in2_fields='speakers', attended_field='label', final_batch_size=256,
                                 shuffle_buffer_size=0)
  bd.set_trials(['trial_%04d' % i for i in range(n_trials)], [frames] * n_trials, eeg, att,
                input2_data=torch.cat([s1, s2], dim=1), attention_data=label)
  widths = [128, 256, 640, 1280, 2560, 3840]              # 1, 2, 5, 10, 20, 30 s windows, step W/2
  def run(gram_options=None):
    return infer.leave_one_trial_out_test(bd, regularization_lambda=0.1, window_list=widths,
                                          reduction='first', decoder_type='wta', frame_rate=128.0,
                                          return_details=True, gram_options=gram_options)
  acc, det = run()                                         # warm-up (allocator, handles)
  n_windows = sum(len(det[w]['attention']) for w in widths)
  solver = det['solver']
  del det
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  acc, det = run()
  torch.cuda.synchronize()
  wall = time.perf_counter() - t0
  del det
  # the same with the per-trial Grams through their block-Toeplitz structure (method 5, opt-in)
  from telluride_decoding_b200 import ops
  acc5, det5 = run(ops.gram_options(method=5))
  del det5
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  acc5, det5 = run(ops.gram_options(method=5))
  torch.cuda.synchronize()
  wall5 = time.perf_counter() - t0
  del det5
  return dict(workload='cfg5: %d trials x 7680 frames (60 s @128 Hz), 64 ch x 32 lags, leave-one-trial-out '
                       'ridge (lambda 0.1, %d folds of a 2049 x 2049 system), 2 speakers, windows '
                       '1/2/5/10/20/30 s (step W/2), reduction first, winner-take-all' % (n_trials, n_trials),
              loo_wall_s=wall, frames_per_s=n / wall, folds_per_s=n_trials / wall,
              windows_per_s=2 * n_windows / wall, decisions=n_windows,
              accuracy={str(w): acc[w] for w in widths}, fold_solver=solver,
              structured=dict(loo_wall_s=wall5, accuracy={str(w): acc5[w] for w in widths},
                              what='per-trial Grams through the block-Toeplitz structure (gram_options(method=5))'),
              note='wall clock of infer.leave_one_trial_out_test (per-trial Grams from one launch, all folds from one '
                   'factorisation + fixed-point steps, held-out predictions, correlator, windows, decisions; data resident); the '
                   'reference runs one complete fit per held-out trial (regression.py:197-214)')


# ----------------------------------------------------------------------------------------------
# This framework: headline
# ----------------------------------------------------------------------------------------------

def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=3)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--legs', default='all', help='all | none | comma list of cfg2,wcorr,cfg3,cfg5,fit')
  ap.add_argument('--e2e-steps', type=int, default=3, help='timed end-to-end steps (each a full recording)')
  ap.add_argument('--minutes', type=float, default=60.0, help='development: shorter recording')
  ap.add_argument('--ref-rows', type=int, default=1000, help='reference arm: frames per minibatch')
  ap.add_argument('--ref-batches', type=int, default=1, help='reference arm: minibatches per step')
  ap.add_argument('--cta-group', type=int, default=0)
  ap.add_argument('--chunk-rows', type=int, default=0)
  ap.add_argument('--gram-method', type=int, default=0)
  ap.add_argument('--skip-cpu-baseline', action='store_true')
  args = ap.parse_args()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))

  if args.impl == 'reference':
    run_reference_arm(args, rank)
    return

  import random
  import torch
  import torch.distributed as dist
  from telluride_decoding_b200 import brain_data, brain_model, ops, sharding

  random.seed(SEED)                                  # same (shuffled) file list on every rank
  assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback)'
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  warmup = max(args.warmup, 3)
  peaks = load_peaks()
  cfg = dict(CFG4)
  if args.minutes != 60.0:
    cfg['trials'] = max(1, int(round(args.minutes * 60000 / cfg['frames_per_trial'])))
    cfg['workload'] = 'DEVELOPMENT SLICE (%g min) of ' % args.minutes + cfg['workload']
  fpt, trials, c = cfg['frames_per_trial'], cfg['trials'], cfg['channels']
  pre, post = cfg['pre'], cfg['post']
  frames = fpt * trials
  d = c * (pre + 1 + post)
  names = ['trial_%02d' % i for i in range(trials)]
  lens = [fpt] * trials
  starts = [i * fpt for i in range(trials)]

  # ---- this rank's share of the recording (frames are shared in storage order) ----------------------
  my_first, my_keep = sharding.shard_rows(None, lens, rank, world)
  my_rows = sum(k - f for f, k in zip(my_first, my_keep))
  row_lo, row_hi = sharding.needed_row_range(starts, lens, lens, pre, post, rank, world)
  # pinned host arrays holding just the rows this rank touches (its share + the lag halo)
  x_pin = torch.empty((row_hi - row_lo, c), dtype=torch.float32, pin_memory=True)
  y_pin = torch.empty((row_hi - row_lo, 1), dtype=torch.float32, pin_memory=True)
  fill_rows(cfg, x_pin, y_pin, row_lo, row_hi)

  def new_brain_data():
    return brain_data.TrialBrainData('meg', 'envelope', 1000.0, pre_context=pre, post_context=post,
                                     final_batch_size=cfg['batch'], shuffle_buffer_size=0)

  gopt = ops.gram_options(method=args.gram_method, cta_group=args.cta_group,
                          chunk_rows=args.chunk_rows, timed=True)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- resident arm: the rank's frames already in HBM ----------------------------------------------
  bd_res = new_brain_data()
  bd_res.set_trials(names, lens, x_pin, y_pin, upload_rows=(row_lo, row_hi))
  ds_res = bd_res.create_dataset('train')
  assert ds_res.num_batches * cfg['batch'] == frames
  torch.cuda.synchronize()

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  stats = None
  for _ in range(warmup):
    stats = None                                      # release the previous Gram before the next one
    stats = sharding.sharded_regression_stats(ds_res, gram_options=gopt)
  barrier()
  gram_ms, ar_ms, step_ms = [], [], []
  win0 = time.time()
  wall0 = time.perf_counter()
  for i in range(args.steps):
    stats = None
    if world > 1:
      dist.barrier()
    tm = {}
    stats = sharding.sharded_regression_stats(ds_res, gram_options=gopt, timings=tm)
    tm['allreduce_done'].synchronize()
    step_ms.append(tm['start'].elapsed_time(tm['allreduce_done']))
    ar_ms.append(tm['stats_done'].elapsed_time(tm['allreduce_done']))
    gram_ms.append(ops.last_gram_kernel_ms())
  barrier()
  wall = time.perf_counter() - wall0
  windows = [(win0, time.time())]
  t_res = torch.tensor([sum(step_ms), float(np.mean(ar_ms)), float(np.mean(gram_ms))],
                       dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
  ms_per_step = float(t_res[0].item()) / args.steps
  allreduce_ms = float(t_res[1].item())
  gram_kernel_ms_max = float(t_res[2].item())
  # The rank that finishes its Gram last waits for nobody: its figure is the collective itself.
  t_min = torch.tensor([float(np.mean(ar_ms))], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
  allreduce_ms_min = float(t_min.item())
  n_stat = stats.n

  # ---- end-to-end arm: host buffers in, fitted weights out ------------------------------------------
  # Every step: pinned host arrays -> (this rank's rows + halo) -> device, statistics, allreduce, ONE
  # ridge solve of the (D+1) x (D+1) system in float64 (replicated on every rank), weights -> host.
  solve_ms = []
  def e2e_step():
    bd = new_brain_data()
    bd.set_trials(names, lens, x_pin, y_pin, upload_rows=(row_lo, row_hi))
    st = sharding.sharded_regression_stats(bd.create_dataset('train'), gram_options=gopt)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g_aug, c_aug = st.augmented(True)
    del st
    w = ops.ridge_solve(g_aug, c_aug, n_stat, [0.1], method=0)[0, :, 0]
    b.record()
    w_host = w.cpu()
    solve_ms.append(a.elapsed_time(b))
    return w_host

  e2e = None
  try:
    stats_gram_check = stats.gram                   # kept for the parity check below
    del stats
    e2e_step()                                       # warm-up (allocator, cuSOLVER handles)
    solve_ms.clear()
    barrier()
    e2e_ev = []
    win0 = time.time()
    for i in range(args.e2e_steps):
      if world > 1:
        dist.barrier()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      w_host = e2e_step()
      b.record()
      e2e_ev.append((a, b))
    barrier()
    windows.append((win0, time.time()))
    t_e2e = torch.tensor([sum(a.elapsed_time(b) for a, b in e2e_ev)], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_e2e.item()) / args.e2e_steps
    e2e = dict(value=frames / (e2e_ms * 1e-3), unit='frames/s',
               h2d_bytes_per_step=int((row_hi - row_lo) * (c + 1) * 4),
               d2h_bytes_per_step=int(w_host.numel() * 4), ms_per_step=e2e_ms, steps=args.e2e_steps,
               solve_ms=float(np.mean(solve_ms)),
               what='TrialBrainData.set_trials(pinned host arrays, this rank\'s rows + lag halo) -> '
                    'sharded_regression_stats -> ridge_solve(lambda 0.1, float64 Cholesky of the '
                    '%d x %d system, replicated per rank) -> weights on the host' % (d + 1, d + 1))
  except Exception as exc:   # pylint: disable=broad-except
    e2e = dict(value=None, unit='frames/s', error=repr(exc)[:300])
    stats_gram_check = None
  clocks = sampler.stop(windows) if rank == 0 else None

  if rank != 0:
    if world > 1:
      sharding.close_native_comms()
      dist.destroy_process_group()
    return

  # ---- roofline of the dominant kernel (the tcgen05 Gram, rank 0's launch) --------------------------
  rows0 = my_rows
  f_alg = float(rows0) * d * (d + 1)                 # symmetric-minimal flops of one launch
  kernel_ms = float(np.mean(gram_ms))
  achieved = f_alg / (kernel_ms * 1e-3) / 1e12
  long_step = kernel_ms > 500.0
  # Three MMAs per product either way: kind::f16 (the bf16 / fp16 dense rate) for the default fp16 x 3
  # flavour, kind::tf32 (half that rate) for --gram-method 4.
  f16 = args.gram_method in (0, 6)
  mma_div = 3.0 if f16 else 6.0
  peak = (peaks['bf16_sustained'] if long_step else peaks['bf16']) / mma_div
  traffic = None
  if world == 1 and args.minutes == 60.0:
    traffic = GRAM_DRAM_TRAFFIC_BYTES_CFG4.get(6 if f16 else (4 if args.gram_method == 4 else -1))
  roofline = dict(
      bound='tensor', achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
      traffic=traffic,
      traffic_source='profiles/r02_ncu_launches_summary.txt (ncu dram__bytes_read/write.sum of this command, per launch); '
                     'algorithmic bytes per launch %.2f GB' % ((float(rows0) * c * 4 + float(d) * d * 4) / 1e9),
      kernel=('hk::lagged_gram_halo_kernel<false, true> (fp16 x 3)' if f16 else
              'hk::lagged_gram_halo_kernel<false, false> (3xTF32)' if args.gram_method == 4 else 'ts::lagged_gram_ts_kernel'),
      flops_per_launch=f_alg,
      frames_per_launch=rows0, kernel_ms=kernel_ms, kernel_ms_max_over_ranks=gram_kernel_ms_max,
      peak_source='%s bf16_tflops%s%s / 3 (three MMAs per product): the %s figure, the kernel runs for '
                  '%.2f s' % (peaks['source'], '_sustained' if long_step else '',
                              ' (= the fp16 rate)' if f16 else ' / 2 (tf32 rate)',
                              'sustained' if long_step else 'burst', kernel_ms * 1e-3),
      frac_of_burst_peak=achieved / (peaks['bf16'] / mma_div),
      frac_of_sustained_peak=achieved / (peaks['bf16_sustained'] / mma_div),
      # BASELINE.json's north star quotes the target against the 3xTF32 peak (bf16 / 2 / 3):
      frac_of_3xtf32_burst_peak=achieved / (peaks['bf16'] / 6.0),
      frac_of_3xtf32_sustained_peak=achieved / (peaks['bf16_sustained'] / 6.0),
      limiter='board power: the kernel runs at the 1000 W cap (sw_power_cap, ~1.5 GHz of 1.965) in both flavours; '
              'fp16 x 3 delivers ~1.45x the split-precision flops per joule of 3xTF32 (profiles/r02_gram_fp16x3.md)',
      share_of_step=kernel_ms / ms_per_step)

  # ---- parity of the bench's own numbers against the CPU oracle (untimed) + cpu_baseline ------------
  cpu = parity = None
  if world == 1 and not args.skip_cpu_baseline:
    use_all_cores()
    rows = args.ref_rows
    nb = 3
    dt, nfr, g_cpu = cpu_cfg4_minibatches(nb, rows, keep_gram=True)
    cpu = dict(value=nfr / dt, unit='frames/s', cores=cpu_cores(), kind='port',
               sample='%d minibatches of %d frames of trial 0 (lag design + x.T@x + sum(x) + x.T@y, '
                      'numpy/BLAS float32); rate of the sample, nothing extrapolated' % (nb, rows),
               gram_gflops=2.0 * nfr / dt * (d + 1) ** 2 / 1e9, seconds=dt)
    # the same rows through the kernel: design rows [0, nb*rows) of trial 0
    sx = ds_res.stream('x')
    keep = [nb * rows if st == 0 else 0 for st in sx.seg_start]   # trial 0 = raw rows from 0 (files are shuffled)
    g_gpu = ops.lagged_gram([sx], keep, options=ops.gram_options(method=args.gram_method))
    g64 = g_gpu.cpu().numpy().astype(np.float64)
    want = g_cpu[:d, :d].astype(np.float64)
    parity = dict(gram_rel_fro=float(np.linalg.norm(g64 - want) / np.linalg.norm(want)),
                  rows='design rows [0, %d) of trial 0, all %d x %d entries' % (nb * rows, d, d),
                  against='oracle port (float32 numpy accumulation)', tolerance=1e-5)
    del g_gpu, g64, want, g_cpu
    # lag blocks of the full-recording Gram against float64 autocorrelations of the raw frames
    if stats_gram_check is not None:
      xd = ds_res.stream('x').data.view(trials, fpt, c)
      errs = []
      for (k1, k2) in ((0, 0), (3, 50), (99, 99), (17, 18), (0, 99)):
        want = torch.zeros((c, c), dtype=torch.float64, device=dev)
        hi = fpt - max(k1, k2)
        for i in range(trials):
          want += xd[i, k1:k1 + hi].to(torch.float64).t() @ xd[i, k2:k2 + hi].to(torch.float64)
        got = stats_gram_check[k1 * c:(k1 + 1) * c, k2 * c:(k2 + 1) * c].to(torch.float64)
        errs.append(float(torch.linalg.norm(got - want) / torch.linalg.norm(want)))
      parity['full_recording_lag_blocks_rel_fro_max'] = max(errs)
      parity['full_recording_note'] = ('five 306 x 306 lag blocks of the timed step\'s Gram vs float64 '
                                       'torch products of the raw frames (not the oracle: 3.6 M rows)')
  # ---- the same Gram through the block-Toeplitz structure (method 5, opt-in; untimed region) ----------
  structured = None
  if world == 1 and stats_gram_check is not None and args.legs != 'none':
    try:
      sx = ds_res.stream('x')
      opt5 = ops.gram_options(method=5, timed=True)
      g5 = None
      t5, k5 = [], []
      for _ in range(3):
        g5 = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g5 = ops.lagged_gram([sx], ds_res.seg_rows, options=opt5)
        b.record()
        torch.cuda.synchronize()
        t5.append(a.elapsed_time(b)); k5.append(ops.last_gram_kernel_ms())
      diff = float(torch.linalg.norm((g5 - stats_gram_check).double()) / torch.linalg.norm(stats_gram_check.double()))
      ms5 = float(np.median(t5))
      structured = dict(
          what='ops.lagged_gram(..., method=5): a lagged Gram is block-Toeplitz up to edge terms; only the first '
               'block row (306 x 30600 outputs) is contracted on the tensor cores (same kernel, rectangular mode), '
               'the other blocks follow along the lag diagonals with edge sums over the first / last frames of each trial',
          ms=ms5, block_row_kernel_ms=float(np.median(k5)), frames_per_s=frames / (ms5 * 1e-3),
          speedup_vs_dense_gram=kernel_ms / ms5, rel_fro_vs_dense_gram=diff,
          note='opt-in; the dense kernel stays the default and the roofline reference (F_alg counts the dense contraction)')
      del g5
    except Exception as exc:   # pylint: disable=broad-except
      structured = dict(error=repr(exc)[:300])
  stats_gram_check = None

  # absmax, split, row scales, gram, symmetrize, thin, thin-reduce (+ memsets / library allreduce, not counted)
  n_launch = 7 if f16 else 5
  legs = {}
  want = ['cfg2', 'wcorr', 'cfg3', 'cfg5'] if args.legs == 'all' else \
         ([] if args.legs == 'none' else args.legs.split(','))
  if world == 1 and want:
    torch.cuda.empty_cache()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    for name, fn in (('cfg2', lambda: leg_cfg2(dev, peaks, flush, not args.skip_cpu_baseline)),
                     ('wcorr', lambda: leg_windowed_corr(dev, peaks)),
                     ('cfg3', lambda: leg_cfg3(dev)),
                     ('cfg5', lambda: leg_cfg5(dev))):
      if name in want:
        try:
          legs[name] = fn()
        except Exception as exc:   # pylint: disable=broad-except
          legs[name] = dict(error=repr(exc)[:300])
        torch.cuda.empty_cache()

  par = ('single GPU' if world == 1 else
         'one recording time-sharded x%d (sharding.shard_rows: contiguous frame ranges, cuts on 32-frame '
         'tiles inside trials, lag halo from the same trial), 1 NCCL allreduce of the fp32 Gram\'s upper '
         'triangle (libtdgram.so\'s own communicator) + 1 of the packed f64 thin sums' % world)
  line = dict(
      metric=METRIC, value=frames / (ms_per_step * 1e-3), unit='frames/s', n_gpus=world,
      steps=args.steps, warmup=warmup, ms_per_step=ms_per_step, higher_is_better=True,
      scaling='strong', vs_baseline=None,
      dtype='%s (fp32-equivalent) Gram, f64 thin sums' % ('fp16x3 split' if f16 else 'tf32x3'),
      data='synthetic',
      config=dict(cfg, parallelism=par, D=d,
                  l2='inputs (%.1f GB of raw frames per rank) are far larger than the 126 MB L2; '
                     'no flush needed' % ((row_hi - row_lo) * c * 4 / 1e9)),
      gpu_launches=n_launch * args.steps,
      tflops_alg=float(frames) * d * (d + 1) / (ms_per_step * 1e-3) / 1e12,
      gram_kernel_ms=kernel_ms, allreduce_ms=allreduce_ms if world > 1 else 0.0,
      allreduce_ms_note='stats_done -> allreduce_done: max over ranks (includes waiting for the slowest rank\'s '
                        'Gram); min over ranks = %.2f ms = the collectives themselves (packed upper triangle of '
                        'the Gram + thin sums, pack / unpack / mirror)' % (allreduce_ms_min if world > 1 else 0.0),
      step_wall_ms=wall / args.steps * 1e3, rows_checked=n_stat == frames,
      roofline=roofline, e2e=e2e, clocks=clocks, cpu_baseline=cpu, parity=parity, structured_gram=structured,
      legs=legs)
  print(json.dumps(line))
  if world > 1:
    sharding.close_native_comms()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
