"""Host-side mirror of tchem::IC::IntCoordSet (reference include/tchem/IC/IntCoordSet.hpp:9-45)
over the C ABI. Same method names, argument meaning and error behaviour as the reference;
every tensor argument may carry a leading batch axis (the reference accepts one geometry per
call, source/IC/IntCoordSet.cpp:140-141). torch is used for device memory and streams only.
"""
import ctypes as C
import math

import torch

from . import _capi
from ._capi import lib, check


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class InvDispView:
    """read-only (coefficient, type, atoms, min) record of one term, cf. IntCoord::operator[]"""
    __slots__ = ("coeff", "type", "atoms", "min")

    def __init__(self, rec):
        self.coeff = rec.coeff
        self.type = _capi.TYPE_NAMES[rec.type]
        self.atoms = [rec.atoms[i] for i in range(rec.natoms)]
        self.min = rec.min


class IntCoordSet:
    """IntCoordSet(format, file): format "Columbus7" or anything else for the default format."""

    def __init__(self, format, file=None, *, text=None, n_atoms=0, device=None):
        n_terms, n_ic = C.c_int(0), C.c_int(0)
        fmt = format.encode()
        if text is not None:
            parse = lambda buf, cap: lib.tchem_ic_parse_text(fmt, text.encode(), buf, cap, C.byref(n_terms), C.byref(n_ic))
        else:
            parse = lambda buf, cap: lib.tchem_ic_parse_file(fmt, str(file).encode(), buf, cap, C.byref(n_terms), C.byref(n_ic))
        check(parse(None, 0))
        self._terms = (_capi.InvDisp * max(n_terms.value, 1))()
        check(parse(self._terms, n_terms.value))
        self._n_terms = n_terms.value
        self._handle = C.c_void_p()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = torch.device("cuda", int(device) if not isinstance(device, torch.device) else (device.index or 0))
        check(lib.tchem_ic_table_create(self._terms, self._n_terms, n_ic.value, int(n_atoms), self.device.index,
                                        C.byref(self._handle)))
        self.intdim = lib.tchem_ic_table_intdim(self._handle)
        self.cartdim = lib.tchem_ic_table_cartdim(self._handle)

    @classmethod
    def from_text(cls, format, text, **kw):
        return cls(format, text=text, **kw)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            lib.tchem_ic_table_destroy(h)
            self._handle = None

    # ---- structure accessors -------------------------------------------------------------
    def size(self):
        return self.intdim

    __len__ = size

    def intcoords(self):
        """list (one entry per internal coordinate) of lists of InvDispView"""
        out = [[] for _ in range(self.intdim)]
        for i in range(self._n_terms):
            out[self._terms[i].ic].append(InvDispView(self._terms[i]))
        return out

    def __getitem__(self, index):
        return self.intcoords()[index]

    # ---- argument plumbing -----------------------------------------------------------------
    def _prep(self, r, name):
        if not isinstance(r, torch.Tensor):
            raise ValueError("tchem::IC::IntCoordSet::%s: r must be a tensor" % name)
        if r.dim() not in (1, 2):
            raise ValueError("tchem::IC::IntCoordSet::%s: r must be a vector (or a batch of vectors)" % name)
        if r.dtype != torch.float64:
            raise ValueError("tchem::IC::IntCoordSet::%s: r must be float64" % name)
        if r.shape[-1] != self.cartdim:
            raise ValueError("tchem::IC::IntCoordSet::%s: r must have %d elements per geometry" % (name, self.cartdim))
        batched = r.dim() == 2
        r2 = (r if batched else r.unsqueeze(0)).contiguous()
        return r2, batched

    def _same(self, t, like, shape, name, what):
        if not isinstance(t, torch.Tensor) or t.dtype != torch.float64:
            raise ValueError("tchem::IC::IntCoordSet::%s: %s must be a float64 tensor" % (name, what))
        if tuple(t.shape) != tuple(shape) and tuple(t.shape) != tuple(shape[1:]):
            raise ValueError("tchem::IC::IntCoordSet::%s: %s has the wrong shape %s" % (name, what, tuple(t.shape)))
        return t.reshape(shape).to(like.device).contiguous()

    def _on_device(self, t):
        if t.is_cuda:
            if t.device != self.device:
                raise ValueError("tchem::IC::IntCoordSet: tensor lives on %s, the table on %s" % (t.device, self.device))
            return t
        return t.to(self.device)

    def _eval(self, r, want_J, want_K, value_path, name, out=None):
        r2, batched = self._prep(r, name)
        B, n, cd = r2.shape[0], self.intdim, self.cartdim
        host = not r2.is_cuda
        if host:
            # results land "on the same device as r" like r.new_empty in the reference; pinned so the
            # download is a direct DMA
            pin = torch.cuda.is_available()
            mk = lambda *s: torch.empty(s, dtype=torch.float64, pin_memory=pin)
        else:
            if r2.device != self.device:
                raise ValueError("tchem::IC::IntCoordSet::%s: r lives on %s, the table on %s" % (name, r2.device, self.device))
            mk = lambda *s: torch.empty(s, dtype=torch.float64, device=self.device)
        if out is not None:
            # caller-provided result tensors (same device as r, contiguous float64, exact shapes)
            shapes = [(B, n)] + ([(B, n, cd)] if want_J else []) + ([(B, n, cd, cd)] if want_K else [])
            if len(out) != len(shapes):
                raise ValueError("tchem::IC::IntCoordSet::%s: out must hold %d tensors" % (name, len(shapes)))
            for o, shp in zip(out, shapes):
                if o.dtype != torch.float64 or not o.is_contiguous() or o.numel() != math.prod(shp) or o.is_cuda == host:
                    raise ValueError("tchem::IC::IntCoordSet::%s: bad out tensor" % name)
            mk_it = iter([o.view(shp) for o, shp in zip(out, shapes)])
            mk = lambda *s: next(mk_it)
        q = mk(B, n)
        J = mk(B, n, cd) if want_J else None
        K = mk(B, n, cd, cd) if want_K else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None
        if B > 0:
            if host:
                check(lib.tchem_ic_eval_host(self._handle, ptr(r2), B, ptr(q), ptr(J), ptr(K), int(value_path)))
            else:
                check(lib.tchem_ic_eval(self._handle, ptr(r2), B, ptr(q), ptr(J), ptr(K), int(value_path),
                                        _stream_ptr(self.device)))
        outs = [q] + ([J] if want_J else []) + ([K] if want_K else [])
        if not batched:
            outs = [o[0] for o in outs]
        return outs

    # ---- q, J, K (source/IC/IntCoordSet.cpp:139-175) ---------------------------------------
    def __call__(self, r):
        return self._eval(r, False, False, True, "operator()")[0]

    def compute_IC_J(self, r, out=None):
        return tuple(self._eval(r, True, False, False, "compute_IC_J", out))

    def compute_IC_J_K(self, r, out=None):
        return tuple(self._eval(r, True, True, False, "compute_IC_J_K", out))

    # ---- gradient / Hessian transforms (source/IC/IntCoordSet.cpp:178-266) -----------------
    #: cart2int transforms factor J J^T per geometry; the reference's at::cholesky raises when it is not positive
    #: definite (source/IC/IntCoordSet.cpp:182). True: read the device flag back after the launch (one
    #: synchronisation) and raise RuntimeError like the reference; False keeps the call asynchronous (the affected
    #: geometries' results are NaN either way, and `factor_status()` can be polled later).
    check_factorisation = True

    def factor_status(self):
        """1 when a cart2int call since the last poll met a non-positive-definite J J^T (synchronises)"""
        return int(lib.tchem_ic_factor_status(self._handle))

    def _transform(self, fn, name, r, extras, out_shape, factors=False, out=None):
        """out: caller-provided result tensor - on the device for device inputs; for host inputs a host tensor
        (pinned: the download is a direct DMA) that the result is copied into before the call returns"""
        r2, batched = self._prep(r, name)
        B = r2.shape[0]
        host = not r2.is_cuda
        dev_r = self._on_device(r2)
        args = []
        for t, shape, what in extras:
            args.append(self._same(t, dev_r, (B,) + shape, name, what))
        if out is not None and (out.dtype != torch.float64 or not out.is_contiguous() or out.numel() != B * math.prod(out_shape)
                                or out.is_cuda == host):
            raise ValueError("tchem::IC::IntCoordSet::%s: bad out tensor" % name)
        res = out.view((B,) + out_shape) if out is not None and not host else \
            torch.empty((B,) + out_shape, dtype=torch.float64, device=self.device)
        ptr = lambda t: C.c_void_p(t.data_ptr())
        if B > 0:
            check(fn(self._handle, ptr(dev_r), *[ptr(a) for a in args], B, ptr(res), _stream_ptr(self.device)))
            if factors and self.check_factorisation and self.factor_status():
                raise RuntimeError("tchem::IC::IntCoordSet::%s: cholesky: J J^T is not positive definite for at least one "
                                   "geometry of the batch (redundant or singular internal coordinates)" % name)
        if host:
            if out is not None:
                hres = out.view((B,) + out_shape)
                hres.copy_(res, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
                res = hres
            else:
                res = res.cpu()
        return res if batched else res[0]

    def gradient_cart2int_matrix(self, r, out=None):
        return self._transform(lib.tchem_ic_grad_cart2int_matrix, "gradient_cart2int", r, [],
                               (self.intdim, self.cartdim), factors=True, out=out)

    def gradient_cart2int(self, r, cartgrad, out=None):
        return self._transform(lib.tchem_ic_grad_cart2int, "gradient_cart2int", r,
                               [(cartgrad, (self.cartdim,), "Cartesian coordinate gradient")], (self.intdim,), factors=True, out=out)

    def gradient_int2cart(self, r, intgrad, out=None):
        return self._transform(lib.tchem_ic_grad_int2cart, "gradient_int2cart", r,
                               [(intgrad, (self.intdim,), "Internal coordinate gradient")], (self.cartdim,), out=out)

    def Hessian_cart2int(self, r, cartgrad, cartHess, out=None):
        return self._transform(lib.tchem_ic_hess_cart2int, "Hessian_cart2int", r,
                               [(cartgrad, (self.cartdim,), "Cartesian coordinate gradient"),
                                (cartHess, (self.cartdim, self.cartdim), "Cartesian coordinate Hessian")],
                               (self.intdim, self.intdim), factors=True, out=out)

    def Hessian_int2cart(self, r, intgrad, intHess, out=None):
        return self._transform(lib.tchem_ic_hess_int2cart, "Hessian_int2cart", r,
                               [(intgrad, (self.intdim,), "Internal coordinate gradient"),
                                (intHess, (self.intdim, self.intdim), "Internal coordinate Hessian")],
                               (self.cartdim, self.cartdim), out=out)
