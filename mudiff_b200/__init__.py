"""Importable alias of the `mu-diff_b200/` package (a hyphen is not a valid identifier):
`import mudiff_b200` gives the very same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = importlib.import_module('mu-diff_b200')
sys.modules[__name__] = _real
