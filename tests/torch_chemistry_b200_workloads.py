"""Loads torch_chemistry_b200/workloads.py without importing the package (whose __init__ needs
the built CUDA library) so that oracle-only tests run anywhere."""
import importlib.util
import os
import sys

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torch_chemistry_b200", "workloads.py")
_spec = importlib.util.spec_from_file_location("_tchem_workloads", _p)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["_tchem_workloads"] = _mod
_spec.loader.exec_module(_mod)
get = _mod.get
CONFIGS = _mod.CONFIGS
