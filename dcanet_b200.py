"""Importable alias: the package directory name mandated for this repo contains hyphens, so
`import dcanet_b200` loads `cost-volume-aggregation-in-stereo-matching-revisited_b200/`."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("cost-volume-aggregation-in-stereo-matching-revisited_b200")
sys.modules[__name__] = _pkg
