"""Development aid: run only td_ridge_solve on a random SPD system (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from telluride_decoding_b200 import ops
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2049
method = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.randn(4 * n, n, device='cuda', dtype=torch.float64)
G = g.t() @ g
C = torch.randn(n, 1, device='cuda', dtype=torch.float64)
lam = np.logspace(-6, 3, 20)
for _ in range(3):
  w = ops.ridge_solve(G, C, 4.0 * n, lam, method=method)
torch.cuda.synchronize()
print('done', float(w.abs().sum()))
