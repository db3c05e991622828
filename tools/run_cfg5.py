"""Development: the cfg5 bench leg alone."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
print(json.dumps(bench.leg_cfg5(torch.device('cuda', 0), n), indent=1))
