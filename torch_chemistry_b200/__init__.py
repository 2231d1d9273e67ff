"""torch_chemistry_b200 - B200-native batched internal coordinates for Torch-Chemistry.

The product is the C-ABI CUDA library (include/tchem_b200.h, csrc/); this package is the thin
Python host binding over it (IntCoordSet / SAPSet with the reference's method names) used by
the tests and bench.py. The C++ drop-in classes live in include/tchem/ + host/.
"""
from ._capi import lib as _lib, LIB_PATH, FileError  # noqa: F401  (import fails if the .so is missing)
from .intcoord import IntCoordSet  # noqa: F401
from .sap import SAPSet  # noqa: F401
from .polynomial import Polynomial, PolynomialSet  # noqa: F401
from .normal_mode import CartNormalMode, IntNormalMode, SANormalMode  # noqa: F401
from . import workloads, sharding  # noqa: F401


def device_count():
    return _lib.tchem_device_count()


def launch_count():
    return _lib.tchem_launch_count()


def kernel_log():
    """{kernel name: launches so far in this process} - what the library actually launched"""
    import ctypes as C
    n = _lib.tchem_kernel_log(None, 0)
    buf = C.create_string_buffer(int(n))
    _lib.tchem_kernel_log(buf, n)
    out = {}
    for item in buf.value.decode().split(";"):
        if item:
            k, v = item.rsplit("=", 1)
            out[k] = int(v)
    return out


def jit_stats():
    """NVRTC bookkeeping: kernels compiled / taken from the disk cache / failed, seconds spent compiling"""
    import ctypes as C
    a = (C.c_int64 * 4)()
    _lib.tchem_jit_stats(a)
    return {"compiled": a[0], "cache_hits": a[1], "failed": a[2], "compile_s": a[3] * 1e-6}
