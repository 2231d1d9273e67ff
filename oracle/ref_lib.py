"""TEST INFRASTRUCTURE ONLY - ctypes binding to oracle/_ref/libtchem_ref.so.

The .so is the reference's own LibTorch CPU implementation (built by oracle/Makefile
from /root/reference). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product
package never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libtchem_ref.so")

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (loads libtorch/libc10 so the .so resolves without rpath luck)
        L = C.CDLL(_SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_ic_create.restype = C.c_void_p
        L.ref_ic_create.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_ic_destroy.argtypes = [C.c_void_p]
        L.ref_ic_intdim.argtypes = [C.c_void_p]
        L.ref_ic_nterms.argtypes = [C.c_void_p]
        L.ref_ic_terms.argtypes = [C.c_void_p, _ip, _dp, _ip, _ip, _ip, _dp]
        L.ref_sap_create.restype = C.c_void_p
        L.ref_sap_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.c_int]
        L.ref_sap_destroy.argtypes = [C.c_void_p]
        L.ref_sap_nterms.argtypes = [C.c_void_p]
        L.ref_set_num_threads.argtypes = [C.c_int]
        for name, nptr in [("ref_ic_q", 2), ("ref_ic_qJ", 3), ("ref_ic_qJK", 4)]:
            f = getattr(L, name)
            f.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64] + [_dp] * (nptr - 1)
        L.ref_ic_grad_int2cart.argtypes = [C.c_void_p, _dp, _dp, C.c_int64, C.c_int64, _dp]
        L.ref_ic_grad_cart2int.argtypes = [C.c_void_p, _dp, _dp, C.c_int64, C.c_int64, _dp]
        L.ref_ic_grad_cart2int_matrix.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int64, _dp]
        L.ref_ic_hess_int2cart.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int64, C.c_int64, _dp]
        L.ref_ic_hess_cart2int.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int64, C.c_int64, _dp]
        L.ref_poly_create.restype = C.c_void_p
        L.ref_poly_create.argtypes = [C.c_int, C.c_int]
        L.ref_poly_destroy.argtypes = [C.c_void_p]
        L.ref_poly_nterms.argtypes = [C.c_void_p]
        L.ref_poly_terms.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.ref_poly_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("ref_sap_value", "ref_sap_jacobian", "ref_sap_jacobian_pow", "ref_sap_jacobian2nd"):
            getattr(L, name).argtypes = [C.c_void_p, _dp, C.c_int64, _dp]
        L.ref_set_num_threads(1)
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _in(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _check(rc):
    if rc != 0:
        raise RuntimeError("reference raised: " + lib().ref_last_error().decode())


class RefIntCoordSet:
    """The reference's tchem::IC::IntCoordSet, called geometry by geometry."""

    def __init__(self, fmt, path):
        self._h = lib().ref_ic_create(fmt.encode(), str(path).encode())
        if not self._h:
            raise RuntimeError("reference raised: " + lib().ref_last_error().decode())
        self.intdim = lib().ref_ic_intdim(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_ic_destroy(self._h)
            self._h = None

    def terms(self):
        """Flattened definition table as parsed (and normalised) by the reference."""
        n = lib().ref_ic_nterms(self._h)
        ic = np.zeros(n, np.int32); coeff = np.zeros(n); typ = np.zeros(n, np.int32)
        nat = np.zeros(n, np.int32); atoms = np.zeros((n, 5), np.int32); mn = np.zeros(n)
        lib().ref_ic_terms(self._h, ic.ctypes.data_as(_ip), _p(coeff), typ.ctypes.data_as(_ip),
                           nat.ctypes.data_as(_ip), atoms.ctypes.data_as(_ip), _p(mn))
        return dict(ic=ic, coeff=coeff, type=typ, natoms=nat, atoms=atoms, min=mn)

    def _r(self, r):
        r = _in(r)
        if r.ndim == 1:
            r = r[None]
        return r, r.shape[0], r.shape[1]

    def q(self, r):
        r, B, cd = self._r(r)
        q = np.empty((B, self.intdim))
        _check(lib().ref_ic_q(self._h, _p(r), B, cd, _p(q)))
        return q

    def qJ(self, r):
        r, B, cd = self._r(r)
        q = np.empty((B, self.intdim)); J = np.empty((B, self.intdim, cd))
        _check(lib().ref_ic_qJ(self._h, _p(r), B, cd, _p(q), _p(J)))
        return q, J

    def qJK(self, r):
        r, B, cd = self._r(r)
        q = np.empty((B, self.intdim)); J = np.empty((B, self.intdim, cd))
        K = np.empty((B, self.intdim, cd, cd))
        _check(lib().ref_ic_qJK(self._h, _p(r), B, cd, _p(q), _p(J), _p(K)))
        return q, J, K

    def gradient_int2cart(self, r, intgrad):
        r, B, cd = self._r(r)
        g = _in(intgrad, (B, self.intdim)); out = np.empty((B, cd))
        _check(lib().ref_ic_grad_int2cart(self._h, _p(r), _p(g), B, cd, _p(out)))
        return out

    def gradient_cart2int(self, r, cartgrad):
        r, B, cd = self._r(r)
        g = _in(cartgrad, (B, cd)); out = np.empty((B, self.intdim))
        _check(lib().ref_ic_grad_cart2int(self._h, _p(r), _p(g), B, cd, _p(out)))
        return out

    def gradient_cart2int_matrix(self, r):
        r, B, cd = self._r(r)
        out = np.empty((B, self.intdim, cd))
        _check(lib().ref_ic_grad_cart2int_matrix(self._h, _p(r), B, cd, _p(out)))
        return out

    def Hessian_int2cart(self, r, intgrad, intHess):
        r, B, cd = self._r(r)
        g = _in(intgrad, (B, self.intdim)); H = _in(intHess, (B, self.intdim, self.intdim))
        out = np.empty((B, cd, cd))
        _check(lib().ref_ic_hess_int2cart(self._h, _p(r), _p(g), _p(H), B, cd, _p(out)))
        return out

    def Hessian_cart2int(self, r, cartgrad, cartHess):
        r, B, cd = self._r(r)
        g = _in(cartgrad, (B, cd)); H = _in(cartHess, (B, cd, cd))
        out = np.empty((B, self.intdim, self.intdim))
        _check(lib().ref_ic_hess_cart2int(self._h, _p(r), _p(g), _p(H), B, cd, _p(out)))
        return out


class RefSAPSet:
    """The reference's tchem::polynomial::SAPSet; xs passed concatenated as x[B, sum(dims)]."""

    def __init__(self, path, irreducible, dims):
        d = (C.c_int64 * len(dims))(*dims)
        self._h = lib().ref_sap_create(str(path).encode(), int(irreducible), d, len(dims))
        if not self._h:
            raise RuntimeError("reference raised: " + lib().ref_last_error().decode())
        self.dims = list(dims)
        self.sumdim = int(sum(dims))
        self.nterms = lib().ref_sap_nterms(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_sap_destroy(self._h)
            self._h = None

    def _x(self, x):
        x = _in(x)
        if x.ndim == 1:
            x = x[None]
        assert x.shape[1] == self.sumdim
        return x, x.shape[0]

    def value(self, x):
        x, B = self._x(x)
        y = np.empty((B, self.nterms))
        _check(lib().ref_sap_value(self._h, _p(x), B, _p(y)))
        return y

    def jacobian(self, x, pow_variant=False):
        x, B = self._x(x)
        J = np.empty((B, self.nterms, self.sumdim))
        f = lib().ref_sap_jacobian_pow if pow_variant else lib().ref_sap_jacobian
        _check(f(self._h, _p(x), B, _p(J)))
        return J

    def jacobian2nd(self, x):
        """SAPSet::Jacobian2nd_(xs, K): K[B, nterms, sumdim, sumdim]"""
        x, B = self._x(x)
        K = np.empty((B, self.nterms, self.sumdim, self.sumdim))
        _check(lib().ref_sap_jacobian2nd(self._h, _p(x), B, _p(K)))
        return K


class RefPolynomialSet:
    """The reference's tchem::polynomial::PolynomialSet(dimension, order)."""

    def __init__(self, dimension, order):
        self._h = lib().ref_poly_create(int(dimension), int(order))
        if not self._h:
            raise RuntimeError("reference raised: " + lib().ref_last_error().decode())
        self.dimension, self.order = int(dimension), int(order)
        self.nterms = lib().ref_poly_nterms(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_poly_destroy(self._h)
            self._h = None

    def terms(self):
        n = lib().ref_poly_terms(self._h, None, None)
        tp, cs = (C.c_int32 * (self.nterms + 1))(), (C.c_int32 * max(n, 1))()
        lib().ref_poly_terms(self._h, tp, cs)
        return [[cs[c] for c in range(tp[i], tp[i + 1])] for i in range(self.nterms)]

    def eval(self, x):
        x = _in(x)
        x = x[None] if x.ndim == 1 else x
        B, d, n = x.shape[0], self.dimension, self.nterms
        y, J, K = np.empty((B, n)), np.empty((B, n, d)), np.empty((B, n, d, d))
        _check(lib().ref_poly_eval(self._h, _p(x), B, _p(y), _p(J), _p(K)))
        return y, J, K
