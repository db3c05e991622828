"""Development aid: run only td_ridge_solve on a random SPD system (for ncu / TD_DEBUG=2) and check
the batched sweep against a float64 torch solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from telluride_decoding_b200 import ops
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2049
method = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
g = torch.randn(4 * n, n, device='cuda', dtype=torch.float64)
G = g.t() @ g
C = torch.randn(n, 1, device='cuda', dtype=torch.float64)
n_lam = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lam = np.logspace(-6, 3, 20)[:n_lam]
for _ in range(3):
  w = ops.ridge_solve(G, C, 4.0 * n, lam, method=method)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
  w = ops.ridge_solve(G, C, 4.0 * n, lam, method=method)
e1.record()
torch.cuda.synchronize()
print('ms per sweep', e0.elapsed_time(e1) / 5)
A = G / (4.0 * n)
worst = 0.0
for i, l in enumerate(lam):
  ref = torch.linalg.solve(A + l * torch.eye(n, device='cuda', dtype=torch.float64), C / (4.0 * n))
  got = w[i].double().reshape(ref.shape) if w[i].numel() == ref.numel() else None
  err = float((got - ref).abs().max() / ref.abs().max())
  worst = max(worst, err)
print('worst relative error vs float64 solve', worst, 'shape', tuple(w.shape))
