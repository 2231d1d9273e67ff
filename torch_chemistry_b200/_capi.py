"""ctypes binding of the C ABI declared in include/tchem_b200.h (libtchem_b200.so).

There is no fallback: if the shared library has not been built the import fails, and every
compute entry point fails with RuntimeError when no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtchem_b200.so")

OK, EINVAL, EDOMAIN, EFILE, ECUDA, ENOMEM = range(6)


class InvDisp(C.Structure):
    """struct tchem_invdisp (include/tchem_b200.h)"""
    _fields_ = [("ic", C.c_int32), ("type", C.c_int32), ("natoms", C.c_int32), ("atoms", C.c_int32 * 5),
                ("coeff", C.c_double), ("min", C.c_double)]


TYPE_NAMES = ["dummy", "stretching", "bending", "cosbending", "torsion", "sintorsion", "costorsion",
              "torsion2", "sintorsion2", "costorsion2", "pxtorsion2", "pytorsion2", "OutOfPlane", "sinoop"]

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "torch_chemistry_b200: %s is missing - build it with `python -m torch_chemistry_b200.build` "
        "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_vp, _dp, _i64, _int = C.c_void_p, C.c_void_p, C.c_int64, C.c_int   # data pointers passed as integers
_pp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int)
_i32p = C.POINTER(C.c_int32)


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("tchem_last_error", C.c_char_p)
_sig("tchem_device_count", _int)
_sig("tchem_launch_count", _i64)
_sig("tchem_kernel_log", _i64, C.c_char_p, _i64)
_sig("tchem_jit_stats", _int, C.POINTER(C.c_int64))
_sig("tchem_ic_parse_file", _int, C.c_char_p, C.c_char_p, C.POINTER(InvDisp), _int, _ip, _ip)
_sig("tchem_ic_parse_text", _int, C.c_char_p, C.c_char_p, C.POINTER(InvDisp), _int, _ip, _ip)
_sig("tchem_sap_parse_file", _int, C.c_char_p, _i32p, _int, _i32p, _int, _ip, _ip)
_sig("tchem_sap_parse_text", _int, C.c_char_p, _i32p, _int, _i32p, _int, _ip, _ip)
_sig("tchem_ic_table_create", _int, C.POINTER(InvDisp), _int, _int, _int, _int, _pp)
_sig("tchem_ic_table_destroy", None, _vp)
_sig("tchem_ic_table_intdim", _int, _vp)
_sig("tchem_ic_table_cartdim", _int, _vp)
_sig("tchem_ic_table_device", _int, _vp)
_sig("tchem_sap_table_create", _int, _i32p, _i32p, _int, _int, _i32p, _int, _int, _pp)
_sig("tchem_sap_table_destroy", None, _vp)
_sig("tchem_sap_table_nterms", _int, _vp)
_sig("tchem_sap_table_sumdim", _int, _vp)
_sig("tchem_ic_eval", _int, _vp, _dp, _i64, _dp, _dp, _dp, _int, _vp)
_sig("tchem_ic_eval_host", _int, _vp, _dp, _i64, _dp, _dp, _dp, _int)
_sig("tchem_ic_eval_timed", _int, _vp, _dp, _i64, _dp, _dp, _dp, _int, _int, _vp, C.POINTER(C.c_float))
_sig("tchem_ic_grad_int2cart", _int, _vp, _dp, _dp, _i64, _dp, _vp)
_sig("tchem_ic_grad_cart2int", _int, _vp, _dp, _dp, _i64, _dp, _vp)
_sig("tchem_ic_grad_cart2int_matrix", _int, _vp, _dp, _i64, _dp, _vp)
_sig("tchem_ic_hess_int2cart", _int, _vp, _dp, _dp, _dp, _i64, _dp, _vp)
_sig("tchem_ic_hess_cart2int", _int, _vp, _dp, _dp, _dp, _i64, _dp, _vp)
_sig("tchem_ic_factor_status", _int, _vp)
_sig("tchem_ic_program_export", _int, C.POINTER(InvDisp), _int, _int, _int, _int, C.POINTER(C.c_int64), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp)
_sig("tchem_ic_jit_source", _i64, C.POINTER(InvDisp), _int, _int, _int, _int, C.c_char_p, _i64)
_sig("tchem_ic_jit_source_bulk", _i64, C.POINTER(InvDisp), _int, _int, _int, _int, C.c_char_p, _i64)
_sig("tchem_ic_jit_source_group", _i64, C.POINTER(InvDisp), _int, _int, _int, _int, C.c_char_p, _i64)
_sig("tchem_sap_jit_hess_source", _i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _int, C.POINTER(C.c_int32), _int, C.c_char_p, _i64)
_sig("tchem_sap_jit_source", _i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _int, C.POINTER(C.c_int32), _int, _int, C.c_char_p, _i64)
_sig("tchem_ic_jit_compile_check", _int, C.c_char_p, C.c_char_p, _i64)
_sig("tchem_ic_uses_specialised", _int, _vp, _i64)
_sig("tchem_ic_sap_uses_specialised", _int, _vp, _vp, _i64)
_sig("tchem_sap_eval", _int, _vp, _dp, _i64, _dp, _dp, _vp)
_sig("tchem_sap_jacobian2nd", _int, _vp, _dp, _i64, _dp, _vp)
_sig("tchem_ic_sap_eval", _int, _vp, _vp, _dp, _i64, _dp, _dp, _dp, _vp)
_sig("tchem_ic_sap_eval_host", _int, _vp, _vp, _dp, _i64, _dp, _dp, _dp)
_sig("tchem_ic_sap_eval_cart", _int, _vp, _vp, _dp, _i64, _dp, _dp, _dp, _vp)
_sig("tchem_ic_sap_jit_cart_source", _i64, C.POINTER(InvDisp), _int, _int, _int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _int, C.POINTER(C.c_int32), _int, C.c_char_p, _i64)
_sig("tchem_ic_sap_jit_cart_bulk_source", _i64, C.POINTER(InvDisp), _int, _int, _int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _int, C.POINTER(C.c_int32), _int, C.c_char_p, _i64)

# every symbol include/tchem_b200.h declares (checked by tests/test_capi.py)
DECLARED = ["tchem_last_error", "tchem_device_count", "tchem_launch_count", "tchem_kernel_log", "tchem_jit_stats", "tchem_ic_parse_file",
            "tchem_ic_parse_text", "tchem_sap_parse_file", "tchem_sap_parse_text", "tchem_ic_table_create",
            "tchem_ic_table_destroy", "tchem_ic_table_intdim", "tchem_ic_table_cartdim", "tchem_ic_table_device",
            "tchem_sap_table_create", "tchem_sap_table_destroy", "tchem_sap_table_nterms", "tchem_sap_table_sumdim",
            "tchem_ic_eval", "tchem_ic_eval_host", "tchem_ic_eval_timed", "tchem_ic_grad_int2cart",
            "tchem_ic_grad_cart2int", "tchem_ic_grad_cart2int_matrix", "tchem_ic_hess_int2cart",
            "tchem_ic_hess_cart2int", "tchem_ic_factor_status", "tchem_ic_program_export", "tchem_ic_jit_source", "tchem_ic_jit_source_bulk", "tchem_ic_jit_source_group", "tchem_sap_jit_hess_source", "tchem_sap_jit_source", "tchem_ic_jit_compile_check", "tchem_ic_uses_specialised", "tchem_ic_sap_uses_specialised", "tchem_sap_eval", "tchem_sap_jacobian2nd", "tchem_ic_sap_eval", "tchem_ic_sap_eval_host", "tchem_ic_sap_eval_cart", "tchem_ic_sap_jit_cart_source", "tchem_ic_sap_jit_cart_bulk_source"]


class FileError(OSError):
    """CL::utility::file_error of the reference (source/IC/IntCoordSet.cpp:16,61)"""


def check(rc):
    """status code -> the exception type the reference's C++ classes throw"""
    if rc == OK:
        return
    msg = lib.tchem_last_error().decode(errors="replace")
    if rc == EINVAL:
        raise ValueError(msg)            # std::invalid_argument
    if rc == EDOMAIN:
        raise ArithmeticError(msg)       # std::domain_error
    if rc == EFILE:
        raise FileError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)              # CUDA errors
