"""Multi-GPU parity check (run under torchrun): frame-sharded statistics + allreduce equal the
single-GPU statistics, and the fitted weights agree."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from telluride_decoding_b200 import brain_data, brain_model, sharding

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rng = np.random.RandomState(3)
bd = brain_data.TrialBrainData('eeg', 'envelope', 128.0, pre_context=3, post_context=12, final_batch_size=500, shuffle_buffer_size=0)
for i, n in enumerate([9000, 7777, 12001]):
  x = rng.randn(n, 24).astype(np.float32)
  y = (x[:, :1] - x[:, 1:2] + 0.3 * rng.randn(n, 1)).astype(np.float32)
  bd.add_trial('t%d' % i, x, y)
import random; random.seed(0)
bd._cached_file_names = sorted(bd.all_files())     # same file order on every rank
ds = bd.create_dataset('train')
full = brain_model.RegressionStats.from_dataset(ds)
shard = sharding.sharded_regression_stats(ds)
eg = float(torch.linalg.norm(shard.gram.double() - full.gram.double()) / torch.linalg.norm(full.gram.double()))
et = float(torch.linalg.norm(shard.xtv - full.xtv) / torch.linalg.norm(full.xtv))
(w1, b1), = brain_model.solve_regression(full, [0.01]); (w2, b2), = brain_model.solve_regression(shard, [0.01])
ew = float(torch.linalg.norm(w1 - w2) / torch.linalg.norm(w1))
ok = eg < 5e-6 and et < 1e-12 and ew < 1e-4 and shard.n == full.n
print('rank %d/%d: gram %.2e thin %.2e weights %.2e n %d -> %s' % (rank, world, eg, et, ew, shard.n, 'OK' if ok else 'FAIL'))
sharding.close_native_comms()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
