"""Importable alias: the package directory is named ``texasholdempocker-rl_b200`` (with a hyphen, after the
reference repository), which ``import`` cannot spell.  ``import texasholdempocker_rl_b200 as phe`` loads it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("texasholdempocker-rl_b200")
sys.modules[__name__] = _pkg
