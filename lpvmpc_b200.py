"""Import shim: the package directory is named ``autonomous-racing-lpv-mpp-mpc_b200`` (not a Python
identifier); ``import lpvmpc_b200`` gives the same module."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("autonomous-racing-lpv-mpp-mpc_b200")
sys.modules[__name__] = _pkg
