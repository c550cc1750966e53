"""`advoc` -- the reference's package name, bound to the B200 implementation.

`import advoc.spectral`, `advoc.audioio`, `advoc.loader` resolve to the modules of `advoc_b200` (same
function names, shapes, dtypes and exceptions as paarthneekhara/advoc's `advoc/` package; SURVEY.md
section 8(b)), so a script written against the reference runs unchanged when this repository is on
the path.  The model classes of the reference's `models/advoc/` live in `advoc_b200.model`."""
import importlib
import sys

for _name in ('spectral', 'audioio', 'loader'):
  _mod = importlib.import_module('advoc_b200.' + _name)
  sys.modules[__name__ + '.' + _name] = _mod
  globals()[_name] = _mod
del _name, _mod
