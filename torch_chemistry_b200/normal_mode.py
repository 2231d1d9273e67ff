"""Batched normal-mode analysis on the GPU: host-side mirror of tchem::chem::CartNormalMode,
IntNormalMode and SANormalMode (reference include/tchem/chem/normal_mode.hpp, source/chem/normal_mode.cpp:23-222).

The natural consumer of the batched Wilson-B matrix and of Hessian_int2cart: every tensor may carry a leading batch axis
(the reference analyses one geometry per object). IntNormalMode is the GF method: G = J M^-1 J^T and the generalized
symmetric eigenproblem G H x = lambda x that the reference hands to LAPACK dsygv (itype 3) through a Fortran shim
(source/FORTRAN/My_dsygv.f90). Here it is the same reduction dsygv performs - Cholesky G = L L^T, C = L^T H L, eigenpairs
(lambda, y) of C, x = L y - as BATCHED device operations: the two reductions are cuBLAS batched products, the factorisation
and the symmetric eigen-solve are cuSOLVER's batched routines (torch.linalg.cholesky / eigh). This layer is library
composition, not a hand-written kernel; the hand-written part of the chain is what produces its inputs (J, H_r).
There is no CPU path: tensors are moved to the CUDA device and the call fails without one."""
import torch


def _device(device):
    if not torch.cuda.is_available():
        raise RuntimeError("tchem::chem: no CUDA device (this library has no CPU path)")
    return torch.device("cuda", torch.cuda.current_device() if device is None else int(device))


def _signed_sqrt(x):
    # frequency^2 -> frequency, negative if imaginary (normal_mode.cpp:49-53, :118-122)
    return torch.where(x > 0, x.abs().sqrt(), -x.abs().sqrt())


class CartNormalMode:
    """CartNormalMode(masses, Hessian): Hessian[..., 3N, 3N] in Cartesian coordinates (normal_mode.cpp:9-76)"""

    def __init__(self, masses, Hessian, device=None):
        H = torch.as_tensor(Hessian, dtype=torch.float64)
        if H.dim() not in (2, 3):
            raise ValueError("tchem::chem::CartNormalMode: Hessian must be a matrix")
        if H.shape[-1] != H.shape[-2]:
            raise ValueError("tchem::chem::CartNormalMode: Hessian must be a square matrix")
        if 3 * len(masses) != H.shape[-1]:
            raise ValueError("tchem::chem::CartNormalMode: inconsistent dimension between masses and Hessian")
        self._host = not H.is_cuda
        self._dev = H.device if H.is_cuda else _device(device)
        self._batched = H.dim() == 3
        self.masses_ = [float(m) for m in masses]
        self.Hessian_ = (H if self._batched else H.unsqueeze(0)).to(self._dev).clone()
        self.ready_ = False

    def _out(self, t):
        t = t if self._batched else t[0]
        return t.cpu() if self._host else t

    def kernel(self):
        s = torch.tensor(self.masses_, dtype=torch.float64, device=self._dev).repeat_interleave(3).sqrt()
        H = self.Hessian_ / s[None, :, None] / s[None, None, :]
        w, v = torch.linalg.eigh(H)
        modes = v.transpose(1, 2)                                   # eigenvectors in rows
        # rule out the 6 modes of smallest |eigenvalue| (translations and rotations), then ascending frequency
        keep = torch.argsort(w.abs(), dim=1, stable=True)[:, 6:]
        freq = _signed_sqrt(torch.gather(w, 1, keep))
        modes = torch.gather(modes, 1, keep[:, :, None].expand(-1, -1, modes.shape[2])) / s[None, None, :]
        order = torch.argsort(freq, dim=1, stable=True)
        self.frequency_ = torch.gather(freq, 1, order)
        self.cartmode_ = torch.gather(modes, 1, order[:, :, None].expand(-1, -1, modes.shape[2]))
        self.ready_ = True

    def _need(self, what):
        if not self.ready_:
            raise RuntimeError("tchem::chem::%s: not ready (call kernel() first)" % what)

    def frequency(self):
        self._need("CartNormalMode::frequency")
        return self._out(self.frequency_)

    def cartmode(self):
        self._need("CartNormalMode::cartmode")
        return self._out(self.cartmode_)


class IntNormalMode(CartNormalMode):
    """IntNormalMode(masses, Jacobian, Hessian): Jacobian[..., n, 3N] of the internal coordinates, Hessian[..., n, n] in
    internal coordinates (normal_mode.cpp:92-150)"""

    def __init__(self, masses, Jacobian, Hessian, device=None):
        J = torch.as_tensor(Jacobian, dtype=torch.float64)
        H = torch.as_tensor(Hessian, dtype=torch.float64)
        if J.dim() not in (2, 3):
            raise ValueError("tchem::chem::IntNormalMode: Jacobian must be a matrix")
        if 3 * len(masses) != J.shape[-1]:
            raise ValueError("tchem::chem::IntNormalMode: inconsistent dimension between masses and Jacobian")
        if H.dim() != J.dim():
            raise ValueError("tchem::chem::IntNormalMode: Hessian must be a matrix")
        if H.shape[-1] != H.shape[-2]:
            raise ValueError("tchem::chem::IntNormalMode: Hessian must be a square matrix")
        if J.shape[-2] != H.shape[-1]:
            raise ValueError("tchem::chem::IntNormalMode: inconsistent dimension between Jacobian and Hessian")
        self._host = not J.is_cuda
        self._dev = J.device if J.is_cuda else _device(device)
        self._batched = J.dim() == 3
        self.masses_ = [float(m) for m in masses]
        self.Jacobian_ = (J if self._batched else J.unsqueeze(0)).to(self._dev).clone()
        self.Hessian_ = (H if self._batched else H.unsqueeze(0)).to(self._dev).clone()
        self.ready_ = False

    def kernel(self):
        J, H = self.Jacobian_, self.Hessian_
        minv = 1.0 / torch.tensor(self.masses_, dtype=torch.float64, device=self._dev).repeat_interleave(3)
        G = (J * minv[None, None, :]) @ J.transpose(1, 2)
        # dsygv itype 3: G = L L^T, C = L^T H L, C y = lambda y, x = L y  (x^T G^-1 x = 1)
        L = torch.linalg.cholesky(G)
        C = L.transpose(1, 2) @ H @ L
        C = 0.5 * (C + C.transpose(1, 2))
        w, Y = torch.linalg.eigh(C)
        self.frequency_ = _signed_sqrt(w)
        self.intmode_ = (L @ Y).transpose(1, 2)                     # eigenvectors in rows
        # Linv = intmode G^-1 = Y^T L^-1
        self.Linv_ = torch.cholesky_solve(self.intmode_.transpose(1, 2), L).transpose(1, 2)
        # cartmode = intmode (J J^T)^-1 J
        Lj = torch.linalg.cholesky(J @ J.transpose(1, 2))
        self.cartmode_ = torch.cholesky_solve(self.intmode_.transpose(1, 2), Lj).transpose(1, 2) @ J
        self.ready_ = True

    def intmode(self):
        self._need("IntNormalMode::intmode")
        return self._out(self.intmode_)

    def Linv(self):
        self._need("IntNormalMode::Linv")
        return self._out(self.Linv_)


class SANormalMode:
    """SANormalMode(masses, Jacobians, Hessians): one (Jacobian, Hessian) pair per irreducible (normal_mode.cpp:156-243)"""

    def __init__(self, masses, Jacobians, Hessians, device=None):
        if len(Jacobians) != len(Hessians):
            raise ValueError("tchem::chem::SANormalMode: inconsistent number of irreducibles between Jacobian and Hessian")
        self._parts = [IntNormalMode(masses, J, H, device=device) for J, H in zip(Jacobians, Hessians)]
        self.ready_ = False

    def kernel(self):
        for p in self._parts:
            p.kernel()
        self.ready_ = True

    def _need(self, what):
        if not self.ready_:
            raise RuntimeError("tchem::chem::SANormalMode::%s: not ready (call kernel() first)" % what)

    def frequencies(self):
        self._need("frequencies")
        return [p.frequency() for p in self._parts]

    def intmodes(self):
        self._need("intmodes")
        return [p.intmode() for p in self._parts]

    def Linvs(self):
        self._need("Linvs")
        return [p.Linv() for p in self._parts]

    def cartmodes(self):
        self._need("cartmodes")
        return [p.cartmode() for p in self._parts]

    def NIrreds(self):
        return len(self._parts)
