"""Host-side mirror of tchem::polynomial::PolynomialSet (reference include/tchem/polynomial/PolynomialSet.hpp:9-80,
source/polynomial/PolynomialSet.cpp:56-179): value, Jacobian and second-order Jacobian of a set of monomials in
`dimension` coordinates. It is the non-symmetry-adapted sibling of SAPSet - one irreducible - and runs on the same
CUDA kernels through the same C ABI (tchem_sap_eval, tchem_sap_jacobian2nd). Batched: x[dimension] or x[B, dimension].
rotation / translation are set-up time utilities and are not provided."""
import itertools

import torch

from .sap import SAPSet


class Polynomial:
    """P(x) = x[coords[0]] * x[coords[1]] * ... (Polynomial.hpp:9-41); coords sorted descending"""

    def __init__(self, coords):
        self._coords = sorted((int(c) for c in coords), reverse=True)

    def coords(self):
        return list(self._coords)

    def order(self):
        return len(self._coords)

    def uniques_orders(self):
        out = {}
        for c in self._coords:
            out[c] = out.get(c, 0) + 1
        return sorted(out.items())


def all_terms(dimension, order):
    """PolynomialSet(dimension, order), source/polynomial/PolynomialSet.cpp:60-111: orders ascending; inside an order the
    dimension-nary counter with former digit >= latter digit (terms compared from the last coordinate to the first)."""
    if dimension == 0:
        return [[]]
    return [list(reversed(c)) for k in range(order + 1) for c in itertools.combinations_with_replacement(range(dimension), k)]


class PolynomialSet:
    def __init__(self, *args, device=None):
        """PolynomialSet(dimension, order) or PolynomialSet(polynomials, dimension) (PolynomialSet.cpp:56-58)"""
        if len(args) != 2:
            raise ValueError("tchem::polynomial::PolynomialSet: (dimension, order) or (polynomials, dimension)")
        if isinstance(args[0], int):
            self._dimension = int(args[0])
            self._polynomials = [Polynomial(t) for t in all_terms(self._dimension, int(args[1]))]
        else:
            self._dimension = int(args[1])
            self._polynomials = [p if isinstance(p, Polynomial) else Polynomial(p) for p in args[0]]
        if self._dimension < 1:
            raise ValueError("tchem::polynomial::PolynomialSet: the CUDA path needs at least one coordinate")
        for p in self._polynomials:
            if any(c < 0 or c >= max(self._dimension, 1) for c in p.coords()):
                raise ValueError("tchem::polynomial::PolynomialSet: coordinate outside the dimension")
        self._max_order = max((p.order() for p in self._polynomials), default=0)
        text = "".join("    ".join("1,%d" % (c + 1) for c in p.coords()) + "\n" for p in self._polynomials)
        self._sap = SAPSet.from_text(text, 0, [max(self._dimension, 1)], device=device)

    def polynomials(self):
        return self._polynomials

    def dimension(self):
        return self._dimension

    def max_order(self):
        return self._max_order

    def orders(self):
        out = [[] for _ in range(self._max_order + 1)]
        for p in self._polynomials:
            out[p.order()].append(p)
        return out

    def __getitem__(self, index):
        return self._polynomials[index]

    def _check(self, x, name):
        if not isinstance(x, torch.Tensor) or x.dim() not in (1, 2):
            raise ValueError("tchem::polynomial::PolynomialSet::%s: x must be a vector" % name)
        if x.shape[-1] != self._dimension:
            raise ValueError("tchem::polynomial::PolynomialSet::%s: x must have a same dimension as the coordinates" % name)
        return [x]

    def __call__(self, x):
        """:133-141"""
        return self._sap(self._check(x, "operator()"))

    def Jacobian_(self, x):
        """[NP, dimension] (:153-161)"""
        return self._sap.Jacobian_cat(self._check(x, "Jacobian_"))

    def Jacobian2nd_(self, x):
        """[NP, dimension, dimension], symmetric (:172-179; Polynomial::Hessian_ mirrors the upper triangle)"""
        return self._sap.Jacobian2nd_cat(self._check(x, "Jacobian2nd_"))

    Jacobian = Jacobian_
    Jacobian2nd = Jacobian2nd_
