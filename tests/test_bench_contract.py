"""bench.py contract pieces that need no GPU: the reference arm (`--impl reference`) prints one
JSON line with the agreed keys, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
  env = dict(os.environ)
  env.update(extra_env or {})
  env.setdefault('OMP_NUM_THREADS', '4')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0', '--gpus', '1', '--ref-rows', '64', '--legs', 'none'], capture_output=True, text=True, env=env, timeout=600)
  assert out.returncode == 0, out.stderr[-2000:]
  return out.stdout.strip()


def test_reference_arm_prints_the_contract_line():
  lines = [l for l in _run().splitlines() if l.startswith('{')]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
              'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
    assert key in d, key
  assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
  assert d['value'] > 0 and d['e2e']['value'] == d['value']
  assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
  assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
  assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_are_silent():
  assert _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == ''
