"""Development: where does the CCA error come from?  Solve from (a) the kernel's statistics and (b) float64
torch statistics, both routes, against the float64 twin of the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_numpy as ora
from telluride_decoding_b200 import cca, ops
from tests import gpu_util

def rel(a, b):
  s = np.sign(np.sum(a * b, axis=0)); return float(np.linalg.norm(a * s - b) / np.linalg.norm(b))

g = np.load(os.path.join('tests', 'golden', 'cca_small.npz'))
n = g['x'].shape[0]
post1, post2, dim, reg, batch = int(g['post1']), int(g['post2']), int(g['dim']), float(g['reg']), int(g['batch'])
bd = gpu_util.make_test_data(g['x'], np.ones((n, 1), np.float32), g['x2'], post_context=post1, in2_post_context=post2, final_batch_size=batch)
ds = bd.create_dataset('train')
batches = gpu_util.oracle_batches(g['x'], np.ones((n, 1), np.float32), g['x2'], None, 0, post1, 0, post2, 0, batch)
twin = ora.cca_from_batches(batches, dim, reg, mini_batch_count=0, dtype=np.float64)
stats = cca.CcaStats.from_dataset(ds)
z = np.concatenate([gpu_util.design64(batches), gpu_util.design64(batches, 'input_2')], axis=1)
g64 = z.T @ z
print('gram rel fro', gpu_util.rel_fro(stats.gram.cpu().numpy(), g64), 'max elementwise rel', float(np.max(np.abs(stats.gram.cpu().numpy() - g64) / np.abs(g64))))
print('sum_x err', float(np.max(np.abs(stats.sum_x.cpu().numpy() - z[:, :stats.d1].sum(0)))), 'frames', stats.frames, z.shape[0])
exact = cca.CcaStats(torch.from_numpy(g64).cuda(), torch.from_numpy(z[:, :stats.d1].sum(0)).cuda(), torch.from_numpy(z[:, stats.d1:].sum(0)).cuda(), float(z.shape[0]), stats.d1)
for name, st in (('kernel stats', stats), ('float64 stats', exact)):
  for eps in (1e-12, 0.02):
    rx, ry, mx, my, e = [t.cpu().numpy() for t in cca.solve_cca(st, dim, reg, eps_eig=eps)]
    print(name, 'eps', eps, 'rot_x', [rel(rx[:, i:i+1], twin[0][:, i:i+1]) for i in range(dim)], 'rot_y', [rel(ry[:, i:i+1], twin[1][:, i:i+1]) for i in range(dim)], 'e', np.abs(e - twin[4]) / twin[4])
print('golden vs twin', [rel(g['rot_x'][:, i:i+1], twin[0][:, i:i+1]) for i in range(dim)])
