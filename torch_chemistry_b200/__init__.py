"""torch_chemistry_b200 - B200-native batched internal coordinates for Torch-Chemistry.

The product is the C-ABI CUDA library (include/tchem_b200.h, csrc/); this package is the thin
Python host binding over it (IntCoordSet / SAPSet with the reference's method names) used by
the tests and bench.py. The C++ drop-in classes live in include/tchem/ + host/.
"""
from ._capi import lib as _lib, LIB_PATH, FileError  # noqa: F401  (import fails if the .so is missing)
from .intcoord import IntCoordSet  # noqa: F401
from .sap import SAPSet  # noqa: F401
from .polynomial import Polynomial, PolynomialSet  # noqa: F401
from . import workloads, sharding  # noqa: F401


def device_count():
    return _lib.tchem_device_count()


def launch_count():
    return _lib.tchem_launch_count()
