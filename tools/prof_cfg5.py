"""Development: cProfile of the cfg 5 leave-one-trial-out pipeline (host view, device synchronised per op)."""
import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['CUDA_LAUNCH_BLOCKING'] = '1'
import torch, bench
from telluride_decoding_b200 import ops
method = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device('cuda', 0)
# build the leg's data by calling the leg once (it returns after its own timing), then profile a second call
import telluride_decoding_b200.infer as infer
orig = infer.leave_one_trial_out_test
calls = []
def wrapped(*a, **k):
  calls.append((a, k))
  return orig(*a, **k)
infer.leave_one_trial_out_test = wrapped
bench.leg_cfg5(dev, 1000)
a, k = calls[0]
k = dict(k); k['gram_options'] = ops.gram_options(method=method)
orig(*a, **k)
pr = cProfile.Profile(); pr.enable()
orig(*a, **k)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:5000])
