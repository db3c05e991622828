"""Empty stand-in for matplotlib so `attention_decoder.py` / `infer_decoder.py` import. TEST INFRASTRUCTURE ONLY."""
import sys
import types


class _Any(types.ModuleType):

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return lambda *a, **k: None


for _n in ('matplotlib.pyplot', 'matplotlib.patches'):
  sys.modules[_n] = _Any(_n)
pyplot = sys.modules['matplotlib.pyplot']
patches = sys.modules['matplotlib.patches']
