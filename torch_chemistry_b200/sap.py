"""Host-side mirror of tchem::polynomial::SAPSet (reference include/tchem/polynomial/SAPSet.hpp:9-86):
value, Jacobian and second-order Jacobian, over the C ABI.

xs is the reference's std::vector<at::Tensor>: one tensor per irreducible, xs[i] of shape
[dims[i]] or batched [B, dims[i]].
"""
import ctypes as C

import torch

from . import _capi
from ._capi import lib, check
from .intcoord import IntCoordSet, _stream_ptr


class SAPSet:
    def __init__(self, sapoly_file, irreducible, dimensions, *, text=None, device=None):
        nt, nc = C.c_int(0), C.c_int(0)
        if text is not None:
            parse = lambda tp, ct, cs, cc: lib.tchem_sap_parse_text(text.encode(), tp, ct, cs, cc, C.byref(nt), C.byref(nc))
        else:
            parse = lambda tp, ct, cs, cc: lib.tchem_sap_parse_file(str(sapoly_file).encode(), tp, ct, cs, cc, C.byref(nt), C.byref(nc))
        check(parse(None, 0, None, 0))
        self._term_ptr = (C.c_int32 * (nt.value + 1))()
        self._coords = (C.c_int32 * max(2 * nc.value, 2))()
        check(parse(self._term_ptr, nt.value, self._coords, nc.value))
        self.irreducible = int(irreducible)
        self.dimensions = [int(d) for d in dimensions]
        dims = (C.c_int32 * len(self.dimensions))(*self.dimensions)
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = torch.device("cuda", int(device) if not isinstance(device, torch.device) else (device.index or 0))
        self._handle = C.c_void_p()
        check(lib.tchem_sap_table_create(self._term_ptr, self._coords, nt.value, self.irreducible, dims,
                                         len(self.dimensions), self.device.index, C.byref(self._handle)))
        self.nterms = lib.tchem_sap_table_nterms(self._handle)
        self.sumdim = lib.tchem_sap_table_sumdim(self._handle)

    @classmethod
    def from_text(cls, text, irreducible, dimensions, **kw):
        return cls(None, irreducible, dimensions, text=text, **kw)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            lib.tchem_sap_table_destroy(h)
            self._handle = None

    def SAPs(self):
        """monomials as lists of (irreducible, index) pairs, sorted descending like SAP::coords_"""
        out = []
        for i in range(self.nterms):
            out.append([(self._coords[2 * c], self._coords[2 * c + 1])
                        for c in range(self._term_ptr[i], self._term_ptr[i + 1])])
        return out

    def _cat(self, xs, name):
        # argument checks of source/polynomial/SAPSet.cpp:135-140
        for x in xs:
            if not isinstance(x, torch.Tensor) or x.dim() not in (1, 2):
                raise ValueError("tchem::polynomial::SAPSet::%s: x must be a vector" % name)
        if len(xs) != len(self.dimensions):
            raise ValueError("tchem::polynomial::SAPSet::%s: x must share a same number of irreducibles as the coordinate system" % name)
        for x, d in zip(xs, self.dimensions):
            if x.shape[-1] != d:
                raise ValueError("tchem::polynomial::SAPSet::%s: x must have a same dimension as the coordinates" % name)
        batched = xs[0].dim() == 2
        x = torch.cat([t if batched else t.unsqueeze(0) for t in xs], dim=1).to(torch.float64)
        was_cuda = x.is_cuda
        return x.to(self.device).contiguous(), batched, was_cuda

    def _run(self, xs, want_y, want_J, name):
        x, batched, was_cuda = self._cat(xs, name)
        B = x.shape[0]
        y = torch.empty((B, self.nterms), dtype=torch.float64, device=self.device) if want_y else None
        J = torch.empty((B, self.nterms, self.sumdim), dtype=torch.float64, device=self.device) if want_J else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None
        if B > 0:
            check(lib.tchem_sap_eval(self._handle, ptr(x), B, ptr(y), ptr(J), _stream_ptr(self.device)))
        outs = []
        for o in (y, J):
            if o is not None:
                o = o if was_cuda else o.cpu()
                outs.append(o if batched else o[0])
        return outs

    def __call__(self, xs):
        """SAPSet::operator(), source/polynomial/SAPSet.cpp:134-144"""
        return self._run(xs, True, False, "operator()")[0]

    def Jacobian_cat(self, xs):
        """the concatenated tensor J that SAPSet::Jacobian_(xs, J) harvests (:178-201)"""
        return self._run(xs, False, True, "Jacobian_")[0]

    def Jacobian_(self, xs):
        """per-irreducible (NSAP x dim_i) blocks, views of the concatenated tensor (:161-201)"""
        J = self.Jacobian_cat(xs)
        return list(torch.split(J, self.dimensions, dim=-1))

    Jacobian = Jacobian_

    def Jacobian2nd_cat(self, xs):
        """the concatenated symmetric tensor K[NSAP, sumdim, sumdim] that SAPSet::Jacobian2nd_(xs, K)
        harvests (source/polynomial/SAPSet.cpp:243-276); batched xs give [B, NSAP, sumdim, sumdim]"""
        x, batched, was_cuda = self._cat(xs, "Jacobian2nd_")
        B = x.shape[0]
        K = torch.empty((B, self.nterms, self.sumdim, self.sumdim), dtype=torch.float64, device=self.device)
        if B > 0:
            check(lib.tchem_sap_jacobian2nd(self._handle, C.c_void_p(x.data_ptr()), B, C.c_void_p(K.data_ptr()),
                                            _stream_ptr(self.device)))
        K = K if was_cuda else K.cpu()
        return K if batched else K[0]

    def Jacobian2nd_(self, xs, with_K=False):
        """CL::utility::matrix of blocks Ks[i][j] (j >= i, None below the diagonal), Ks[i][j] of shape
        [NSAP, dim_i, dim_j]. with_K=False mirrors Jacobian2nd_(xs) (:221-240): diagonal blocks hold
        the upper triangle only; with_K=True mirrors Jacobian2nd_(xs, K) (:243-276) and returns (Ks, K)
        with the blocks as views of the symmetric K."""
        K = self.Jacobian2nd_cat(xs)
        n = len(self.dimensions)
        off = [0]
        for d in self.dimensions:
            off.append(off[-1] + d)
        Ks = [[None] * n for _ in range(n)]
        for i in range(n):
            for j in range(i, n):
                blk = K[..., off[i]:off[i + 1], off[j]:off[j + 1]]
                Ks[i][j] = blk if with_K or i != j else torch.triu(blk)
        return (Ks, K) if with_K else Ks

    Jacobian2nd = Jacobian2nd_

    def value_and_jacobian(self, xs):
        y, J = self._run(xs, True, True, "Jacobian_")
        return y, J

    def eval_cat(self, x, out=None):
        """value and concatenated Jacobian for coordinates already concatenated over the irreducibles:
        x[B, sum(dimensions)] on the table's device -> (y[B, NSAP], J[B, NSAP, sum(dimensions)]); one
        kernel launch, no copies (`out` = caller-provided (y, J))."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.shape[1] == self.sumdim
                and x.dtype == torch.float64 and x.is_contiguous()):
            raise ValueError("tchem::polynomial::SAPSet::eval_cat: x must be a contiguous CUDA float64 [B, %d] tensor" % self.sumdim)
        B = x.shape[0]
        if out is None:
            y = torch.empty((B, self.nterms), dtype=torch.float64, device=self.device)
            J = torch.empty((B, self.nterms, self.sumdim), dtype=torch.float64, device=self.device)
        else:
            y, J = out
        if B > 0:
            check(lib.tchem_sap_eval(self._handle, C.c_void_p(x.data_ptr()), B, C.c_void_p(y.data_ptr()), C.c_void_p(J.data_ptr()),
                                     _stream_ptr(self.device)))
        return y, J

    # ---- chained path ----------------------------------------------------------------------
    def chained(self, intcoordset, r, return_q=False, out=None):
        """r -> q (compute_IC_J formulas) -> SAP values and d SAP / d q in one kernel. Host tensors
        go through the pipelined host-buffer entry point and give host results; `out` = (y, J[, q])
        preallocated result tensors on the same side as r."""
        if not isinstance(intcoordset, IntCoordSet):
            raise ValueError("tchem::polynomial::SAPSet::chained: needs an IntCoordSet")
        r2, batched = intcoordset._prep(r, "compute_IC_J")
        host = not r2.is_cuda
        B = r2.shape[0]
        if out is not None:
            y, J = out[0].view(B, self.nterms), out[1].view(B, self.nterms, self.sumdim)
            q = out[2].view(B, self.sumdim) if return_q else None
        else:
            kw = dict(dtype=torch.float64, pin_memory=True) if host else dict(dtype=torch.float64, device=self.device)
            y = torch.empty((B, self.nterms), **kw)
            J = torch.empty((B, self.nterms, self.sumdim), **kw)
            q = torch.empty((B, self.sumdim), **kw) if return_q else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None
        if B > 0:
            if host:
                check(lib.tchem_ic_sap_eval_host(intcoordset._handle, self._handle, ptr(r2), B, ptr(q), ptr(y), ptr(J)))
            else:
                check(lib.tchem_ic_sap_eval(intcoordset._handle, self._handle, ptr(r2), B, ptr(q), ptr(y), ptr(J),
                                            _stream_ptr(self.device)))
        outs = [y, J] + ([q] if return_q else [])
        return tuple(o if batched else o[0] for o in outs)

    def chained_cart(self, intcoordset, r, second_order=False):
        """r -> (y, dy/dr[, d2y/dr2]): the SAP values and their CARTESIAN derivatives
        dy/dr = (dSAP/dq) J and d2y/dr2 = J^T (d2SAP/dq2) J + sum_k (dSAP/dq_k) K[k] - what callers of the reference
        compose by hand (test/normal_mode/main.cpp:56-61; SAPSet.cpp:178-201, :243-276, IntCoordSet.cpp:147-175).
        First order on large batches of small molecules is one pair-specialised kernel."""
        if not isinstance(intcoordset, IntCoordSet):
            raise ValueError("tchem::polynomial::SAPSet::chained_cart: needs an IntCoordSet")
        r2, batched = intcoordset._prep(r, "compute_IC_J")
        host = not r2.is_cuda
        rd = r2.to(self.device)
        B, nt, cd = rd.shape[0], self.nterms, intcoordset.cartdim
        kw = dict(dtype=torch.float64, device=self.device)
        y = torch.empty((B, nt), **kw)
        g = torch.empty((B, nt, cd), **kw)
        h = torch.empty((B, nt, cd, cd), **kw) if second_order else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None
        if B > 0:
            check(lib.tchem_ic_sap_eval_cart(intcoordset._handle, self._handle, ptr(rd), B, ptr(y), ptr(g), ptr(h),
                                             _stream_ptr(self.device)))
        outs = [y, g] + ([h] if second_order else [])
        if host:
            outs = [o.cpu() for o in outs]
        return tuple(o if batched else o[0] for o in outs)
