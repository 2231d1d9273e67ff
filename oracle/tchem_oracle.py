"""CPU oracle: a numpy restatement of Torch-Chemistry's internal-coordinate hot path.

TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product package never does.

PARITY PINNED: tests/test_oracle.py checks this restatement against
  (a) the reference's own known answers (test/SApolynomial/main.cpp:9-25) and the Columbus
      golden vector test/intcoord/input/intgeom (test/intcoord/main.cpp:9-21),
  (b) golden vectors in tests/golden/ produced by the reference itself (oracle/_ref, built
      from /root/reference; generator tests/golden/make_golden.py), and
  (c) when oracle/_ref is present, directly against it on seeded random batches.

Everything is batched over a leading geometry axis (the reference takes one geometry per
call, source/IC/IntCoordSet.cpp:140-141; looping it over a batch is what this models).
Every function cites the reference lines it restates; all citations are relative to
/root/reference/.

Second derivatives: the reference has no analytic K. It differentiates its analytic J
expression with autograd, one component at a time (source/IC/InvDisp.cpp:859-868), and for
the 5-atom torsion2 family obtains J and K from first/second-order autograd of the value
expression (source/IC/InvDisp.cpp:893-1065). The restatement does the same thing with
forward-mode dual numbers (class Jet below) applied to the same expressions.
"""
import math
import re

import numpy as np

TYPES = ["dummy", "stretching", "bending", "cosbending", "torsion", "sintorsion", "costorsion",
         "torsion2", "sintorsion2", "costorsion2", "pxtorsion2", "pytorsion2", "OutOfPlane", "sinoop"]
# include/tchem/IC/InvDisp.hpp:9-51
NATOMS = {"dummy": 0, "stretching": 2, "bending": 3, "cosbending": 3, "torsion": 4, "sintorsion": 4,
          "costorsion": 4, "torsion2": 5, "sintorsion2": 5, "costorsion2": 5, "pxtorsion2": 5,
          "pytorsion2": 5, "OutOfPlane": 4, "sinoop": 4}
TORSION_LIKE = ("torsion", "torsion2")


# ----------------------------------------------------------------------------------------
# definition-file parsers
# ----------------------------------------------------------------------------------------
class InvDisp:
    """include/tchem/IC/InvDisp.hpp:54-81"""

    def __init__(self, type_, atoms, min_=-math.pi):
        if type_ not in TYPES:  # source/IC/InvDisp.cpp:17
            raise ValueError("Unimplemented internal coordinate type " + type_)
        self.type = type_
        self.atoms = [int(a) for a in atoms]
        self.min = float(min_)


def parse_intcoordset(fmt, text):
    """source/IC/IntCoordSet.cpp:11-120. Returns a list of ICs, each a list of [coeff, InvDisp],
    coefficients normalised (source/IC/IntCoord.cpp:26-31)."""
    ics = []
    lines = text.split("\n")
    # std::getline + `if (!ifs.good()) break`: a last line without a trailing newline is dropped
    if lines and lines[-1] == "":
        lines = lines[:-1]
    elif lines:
        lines = lines[:-1]
    if fmt == "Columbus7":
        for line in lines[1:]:  # first line "TEXAS" skipped (:17)
            if line[:1] == "K":
                ics.append([])
            coeff = 1.0
            if line[10:20] != " " * 10:
                coeff = float(line[10:20])
            key = line[20:24]
            f = lambda a, n: int(float(line[a:a + n])) - 1  # std::stoul on "7." reads 7
            if key == "STRE":
                typ, atoms = "stretching", [f(28, 5), f(34, 9)]
            elif key == "BEND":  # vertex listed last in the file (:31-36)
                typ, atoms = "bending", [f(28, 6), f(45, 9), f(35, 9)]
            elif key == "TORS":
                typ, atoms = "torsion", [f(28, 6), f(35, 9), f(45, 9), f(55, 9)]
            elif key[:3] == "OUT":  # centre listed last (:44-49)
                typ, atoms = "OutOfPlane", [f(28, 6), f(55, 9), f(35, 9), f(45, 9)]
            else:
                break
            ics[-1].append([coeff, InvDisp(typ, atoms)])
    else:
        number = re.compile(r"-?\d+\.?\d*")
        for line in lines:
            strs = line.split()
            if line[:6] != " " * 6:  # :68-71
                ics.append([])
                strs.pop(0)
            coeff = float(strs.pop(0))
            typ = strs.pop(0)
            if typ not in NATOMS:
                raise ValueError("Error during reading internal coordinate definition: "
                                 "unsupported internal coordinate type: " + typ)
            atoms = [int(strs.pop(0)) - 1 for _ in range(NATOMS[typ])]
            mn = -math.pi
            if strs and number.fullmatch(strs[0]):  # :112-113
                mn = float(strs[0])
            ics[-1].append([coeff, InvDisp(typ, atoms, mn)])
    for ic in ics:  # IntCoord::normalize, source/IC/IntCoord.cpp:26-31
        norm2 = 0.0
        for c, _ in ic:
            norm2 += c * c
        norm2 /= math.sqrt(norm2)
        for term in ic:
            term[0] /= norm2
    return ics


def parse_sapset(text):
    """source/polynomial/SAPSet.cpp:89-109 + source/polynomial/SAP.cpp:42-76.
    Returns a list of monomials; each monomial is the list of (irreducible, index) pairs
    sorted descending (SAP.cpp:51). An empty list is the constant term."""
    lines = text.split("\n")
    lines = lines[:-1]  # getline/good() idiom, see parse_intcoordset
    saps = []
    for line in lines:
        strs = line.split()
        if "#" in strs:
            strs = strs[:strs.index("#")]
        coords = []
        for s in strs:
            a, b = [p for p in s.split(",") if p != ""]
            coords.append((int(a) - 1, int(b) - 1))
        coords.sort(reverse=True)
        saps.append(coords)
    return saps


# ----------------------------------------------------------------------------------------
# forward-mode dual numbers over a batch (value, gradient and optionally Hessian w.r.t. n seeds)
# ----------------------------------------------------------------------------------------
class Jet:
    __slots__ = ("v", "g", "h")

    def __init__(self, v, g, h=None):
        self.v, self.g, self.h = v, g, h

    @staticmethod
    def seeds(x, order):
        """x: [B, n] -> list of n Jets, the i-th carrying unit gradient e_i."""
        B, n = x.shape
        out = []
        for i in range(n):
            g = np.zeros((B, n)); g[:, i] = 1.0
            h = np.zeros((B, n, n)) if order == 2 else None
            out.append(Jet(x[:, i].copy(), g, h))
        return out

    def _lift(self, o):
        if isinstance(o, Jet):
            return o
        o = np.broadcast_to(np.asarray(o, dtype=np.float64), self.v.shape)
        return Jet(o, np.zeros_like(self.g), None if self.h is None else np.zeros_like(self.h))

    def __neg__(self):
        return Jet(-self.v, -self.g, None if self.h is None else -self.h)

    def __add__(self, o):
        o = self._lift(o)
        return Jet(self.v + o.v, self.g + o.g, None if self.h is None else self.h + o.h)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o)
        return Jet(self.v - o.v, self.g - o.g, None if self.h is None else self.h - o.h)

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        a, b = self, o
        v = a.v * b.v
        g = a.v[:, None] * b.g + b.v[:, None] * a.g
        h = None
        if a.h is not None:
            ab = a.g[:, :, None] * b.g[:, None, :]
            h = a.v[:, None, None] * b.h + b.v[:, None, None] * a.h + ab + ab.transpose(0, 2, 1)
        return Jet(v, g, h)
    __rmul__ = __mul__

    def recip(self):
        v = 1.0 / self.v
        g = -self.g * (v * v)[:, None]
        h = None
        if self.h is not None:
            gg = self.g[:, :, None] * self.g[:, None, :]
            h = -self.h * (v * v)[:, None, None] + 2.0 * gg * (v * v * v)[:, None, None]
        return Jet(v, g, h)

    def __truediv__(self, o):
        if not isinstance(o, Jet):
            return self * (1.0 / np.asarray(o, dtype=np.float64))
        return self * o.recip()

    def __rtruediv__(self, o):
        return self._lift(o) * self.recip()

    def sqrt(self):
        s = np.sqrt(self.v)
        g = self.g / (2.0 * s)[:, None]
        h = None
        if self.h is not None:
            gg = self.g[:, :, None] * self.g[:, None, :]
            h = self.h / (2.0 * s)[:, None, None] - gg / (4.0 * s * s * s)[:, None, None]
        return Jet(s, g, h)


def _sqrt(x):
    return x.sqrt() if isinstance(x, Jet) else np.sqrt(x)


def _val(x):
    return x.v if isinstance(x, Jet) else x


# 3-vectors are python lists of three scalars (ndarray[B] or Jet)
def _sub(a, b): return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
def _add(a, b): return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]
def _neg(a): return [-a[0], -a[1], -a[2]]
def _scale(s, a): return [s * a[0], s * a[1], s * a[2]]
def _divs(a, s): return [a[0] / s, a[1] / s, a[2] / s]
def _dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def _norm(a): return _sqrt(_dot(a, a))
def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _wrap(phi, mn):
    """[-pi, pi] -> [min, min + 2pi]  (source/IC/InvDisp.cpp:116-118, 346-347)"""
    phi = np.where(phi < mn, phi + 2.0 * math.pi, np.where(phi > mn + 2.0 * math.pi, phi - 2.0 * math.pi, phi))
    return phi


def _acos_signed(cosphi, sign_test):
    """source/IC/InvDisp.cpp:108-115 (and 339-345): clamp, acos, negate when the sign test < 0."""
    with np.errstate(invalid="ignore"):
        phi = np.arccos(np.clip(cosphi, -1.0, 1.0))
    phi = np.where(cosphi > 1.0, 0.0, np.where(cosphi < -1.0, math.pi, phi))
    inside = (cosphi <= 1.0) & (cosphi >= -1.0)
    return np.where(inside & (sign_test < 0.0), -phi, phi)


# ----------------------------------------------------------------------------------------
# InvDisp::operator() - value only, the clamped formulas (source/IC/InvDisp.cpp:64-254)
# ----------------------------------------------------------------------------------------
def invdisp_value(d, r):
    """r: [B, cartdim] -> [B]"""
    B = r.shape[0]
    at = lambda k: [r[:, 3 * d.atoms[k] + c] for c in range(3)]
    t = d.type
    if t == "dummy":  # :68
        return np.ones(B)
    if t == "stretching":  # :69-73
        return _norm(_sub(at(1), at(0)))
    if t in ("bending", "cosbending"):  # :74-97
        r21 = _sub(at(0), at(1)); r21 = _divs(r21, _norm(r21))
        r23 = _sub(at(2), at(1)); r23 = _divs(r23, _norm(r23))
        c = _dot(r21, r23)
        if t == "cosbending":
            return c
        with np.errstate(invalid="ignore"):
            th = np.arccos(np.clip(c, -1.0, 1.0))
        return np.where(c > 1.0, 0.0, np.where(c < -1.0, math.pi, th))
    if t in ("torsion", "sintorsion", "costorsion"):  # :98-142
        r12 = _sub(at(1), at(0)); r23 = _sub(at(2), at(1)); r34 = _sub(at(3), at(2))
        n123 = _cross(r12, r23); n123 = _divs(n123, _norm(n123))
        n234 = _cross(r23, r34); n234 = _divs(n234, _norm(n234))
        if t == "sintorsion":
            return _dot(_cross(n123, n234), _divs(r23, _norm(r23)))
        cosphi = _dot(n123, n234)
        if t == "costorsion":
            return cosphi
        phi = _acos_signed(cosphi, _dot(n123, _cross(n234, r23)))
        return _wrap(phi, d.min)
    if t in ("torsion2", "sintorsion2", "costorsion2", "pxtorsion2", "pytorsion2"):  # :143-225
        r12 = _sub(at(1), at(0)); r23 = _sub(at(2), at(1)); r45 = _sub(at(4), at(3))
        n123 = _cross(r12, r23); n123 = _divs(n123, _norm(n123))
        n2345 = _cross(r23, r45); n2345 = _divs(n2345, _norm(n2345))
        cosphi = _dot(n123, n2345)
        if t == "costorsion2":
            return cosphi
        if t == "torsion2":
            phi = _acos_signed(cosphi, _dot(n123, _cross(n2345, r23)))
            return _wrap(phi, d.min)
        u23 = _divs(r23, _norm(r23))
        sinphi = _dot(_cross(n123, n2345), u23)
        if t == "sintorsion2":
            return sinphi
        u12 = _divs(r12, _norm(r12))
        costheta = -(_dot(u12, u23))
        sintheta = np.sqrt(1.0 - costheta * costheta)
        return sintheta * (cosphi if t == "pxtorsion2" else sinphi)
    if t in ("OutOfPlane", "sinoop"):  # :226-251
        r21 = _sub(at(0), at(1)); r23 = _sub(at(2), at(1)); r24 = _sub(at(3), at(1))
        n234 = _cross(r23, r24); n234 = _divs(n234, _norm(n234))
        s = _dot(n234, _divs(r21, _norm(r21)))
        if t == "sinoop":
            return s
        with np.errstate(invalid="ignore"):
            th = np.arcsin(np.clip(s, -1.0, 1.0))
        return np.where(s < -1.0, -math.pi / 2.0, np.where(s > 1.0, math.pi / 2.0, th))
    raise ValueError(t)


# ----------------------------------------------------------------------------------------
# analytic value + per-atom gradient blocks, the un-clamped "J path" formulas
# (source/IC/InvDisp.cpp:256-634, repeated at :653-852). Generic over ndarray / Jet so that
# K can be obtained by differentiating exactly these expressions.
# ----------------------------------------------------------------------------------------
def _torsion_pieces(a):
    r12 = _sub(a[1], a[0]); n12 = _norm(r12); u12 = _divs(r12, n12)
    r23 = _sub(a[2], a[1]); n23 = _norm(r23); u23 = _divs(r23, n23)
    r34 = _sub(a[3], a[2]); n34 = _norm(r34); u34 = _divs(r34, n34)
    cos123 = -(_dot(u12, u23)); sin123 = _sqrt(1.0 - cos123 * cos123)
    cos234 = -(_dot(u23, u34)); sin234 = _sqrt(1.0 - cos234 * cos234)
    n123 = _divs(_cross(u12, u23), sin123)
    n234 = _divs(_cross(u23, u34), sin234)
    Js = [
        _divs(_neg(n123), n12 * sin123),
        _sub(_scale((n23 - n12 * cos123) / (n12 * n23 * sin123), n123), _scale(cos234 / (n23 * sin234), n234)),
        _add(_scale((n34 * cos234 - n23) / (n23 * n34 * sin234), n234), _scale(cos123 / (n23 * sin123), n123)),
        _divs(n234, n34 * sin234),
    ]
    return u23, n123, n234, Js


def _sintheta_J(a):
    """J of sin(angle 1-2-3) used by px/pytorsion2 (source/IC/InvDisp.cpp:521-535)."""
    r12 = _sub(a[1], a[0]); r23 = _sub(a[2], a[1])
    n21 = _norm(r12); n23 = _norm(r23)
    u21 = _divs(_neg(r12), n21); u23 = _divs(r23, n23)
    c = _dot(u21, u23)
    s = _sqrt(1.0 - c * c)
    J0 = _divs(_scale(c, _sub(_scale(c, u21), u23)), s * n21)
    J2 = _divs(_scale(c, _sub(_scale(c, u23), u21)), s * n23)
    J1 = _sub(_neg(J0), J2)
    return s, [J0, J1, J2]


def _sincos_torsion2(a):
    """sin/cos of the 5-atom dihedral (source/IC/InvDisp.cpp:419-429)."""
    r12 = _sub(a[1], a[0]); r23 = _sub(a[2], a[1]); r45 = _sub(a[4], a[3])
    n123 = _cross(r12, r23); n123 = _divs(n123, _norm(n123))
    n2345 = _cross(r23, r45); n2345 = _divs(n2345, _norm(n2345))
    sinphi = _dot(_cross(n123, n2345), _divs(r23, _norm(r23)))
    cosphi = _dot(n123, n2345)
    sign_test = _dot(n123, _cross(n2345, r23))
    return sinphi, cosphi, sign_test


def _analytic_qJ(t, a, mn):
    """Types whose J is analytic in the reference. a: list of per-atom 3-vectors.
    Returns (q, [per-atom J 3-vectors]); q is ndarray or Jet like the inputs."""
    if t == "stretching":  # :261-272
        r12 = _sub(a[1], a[0]); n = _norm(r12); u = _divs(r12, n)
        return n, [_neg(u), u]
    if t in ("bending", "cosbending"):  # :273-316
        r21 = _sub(a[0], a[1]); n21 = _norm(r21); u21 = _divs(r21, n21)
        r23 = _sub(a[2], a[1]); n23 = _norm(r23); u23 = _divs(r23, n23)
        c = _dot(u21, u23)
        if t == "bending":
            s = _sqrt(1.0 - c * c)
            J0 = _divs(_sub(_scale(c, u21), u23), s * n21)
            J2 = _divs(_sub(_scale(c, u23), u21), s * n23)
            q = np.arccos(_val(c))
        else:
            J0 = _divs(_sub(u23, _scale(c, u21)), n21)
            J2 = _divs(_sub(u21, _scale(c, u23)), n23)
            q = c
        return q, [J0, _sub(_neg(J0), J2), J2]
    if t in ("torsion", "sintorsion", "costorsion"):  # :317-414
        u23, n123, n234, Js = _torsion_pieces(a)
        cosphi = _dot(n123, n234)
        if t == "torsion":
            phi = _acos_signed(_val(cosphi), _val(_dot(n123, _cross(n234, u23))))
            return _wrap(phi, mn), Js
        sinphi = _dot(_cross(n123, n234), u23)
        if t == "sintorsion":
            return sinphi, [_scale(cosphi, j) for j in Js]
        return cosphi, [_scale(-sinphi, j) for j in Js]
    if t in ("OutOfPlane", "sinoop"):  # :571-631
        r21 = _sub(a[0], a[1]); n21 = _norm(r21); u21 = _divs(r21, n21)
        r23 = _sub(a[2], a[1]); n23 = _norm(r23); u23 = _divs(r23, n23)
        r24 = _sub(a[3], a[1]); n24 = _norm(r24); u24 = _divs(r24, n24)
        c324 = _dot(u23, u24)
        s324sq = 1.0 - c324 * c324
        s324 = _sqrt(s324sq)
        st = _dot(u23, _cross(u24, u21)) / s324
        if t == "OutOfPlane":
            ct = _sqrt(1.0 - st * st)
            tt = st / ct
            J0 = _divs(_sub(_divs(_divs(_cross(u23, u24), ct), s324), _scale(tt, u21)), n21)
            J2 = _divs(_sub(_divs(_divs(_cross(u24, u21), ct), s324), _scale(tt / s324sq, _sub(u23, _scale(c324, u24)))), n23)
            J3 = _divs(_sub(_divs(_divs(_cross(u21, u23), ct), s324), _scale(tt / s324sq, _sub(u24, _scale(c324, u23)))), n24)
            q = np.arcsin(_val(st))
        else:
            J0 = _divs(_sub(_divs(_cross(u23, u24), s324), _scale(st, u21)), n21)
            J2 = _divs(_sub(_divs(_cross(u24, u21), s324), _scale(st / s324sq, _sub(u23, _scale(c324, u24)))), n23)
            J3 = _divs(_sub(_divs(_cross(u21, u23), s324), _scale(st / s324sq, _sub(u24, _scale(c324, u23)))), n24)
            q = st
        return q, [J0, _sub(_sub(_neg(J0), J2), J3), J2, J3]
    raise ValueError(t)


def _local_coords(d, r):
    return np.stack([r[:, 3 * a + c] for a in d.atoms for c in range(3)], axis=1)  # [B, 3n]


def _atoms_of(x):
    """x: list of 3n scalars -> list of n 3-vectors"""
    return [x[3 * i:3 * i + 3] for i in range(len(x) // 3)]


def invdisp_qJK(d, r, want_K):
    """InvDisp::compute_IC_J (want_K False) / compute_IC_J_K (want_K True).
    Returns q [B], Jloc [B, 3n] and Kloc [B, 3n, 3n] in the InvDisp's own atom order."""
    B = r.shape[0]
    t = d.type
    n = len(d.atoms)
    if t == "dummy":  # :260, :640
        return np.ones(B), np.zeros((B, 0)), (np.zeros((B, 0, 0)) if want_K else None)
    x = _local_coords(d, r)
    if NATOMS[t] <= 4:
        if not want_K:
            q, Js = _analytic_qJ(t, _atoms_of([x[:, i] for i in range(3 * n)]), d.min)
            return q, np.stack([c for j in Js for c in j], axis=1), None
        # K = derivative of the analytic J expression (source/IC/InvDisp.cpp:859-868), blocks
        # (i, j>=i) taken from d Js[i] / d rs[j] and mirrored (:869-876)
        q, Js = _analytic_qJ(t, _atoms_of(Jet.seeds(x, 1)), d.min)
        Jl = np.stack([c.v for j in Js for c in j], axis=1)
        dJ = np.stack([c.g for j in Js for c in j], axis=1)  # [B, 3n (J comp), 3n (coord)]
        K = np.zeros((B, 3 * n, 3 * n))
        for i in range(n):
            for j in range(i, n):
                blk = dJ[:, 3 * i:3 * i + 3, 3 * j:3 * j + 3]
                K[:, 3 * i:3 * i + 3, 3 * j:3 * j + 3] = blk
                if j > i:
                    K[:, 3 * j:3 * j + 3, 3 * i:3 * i + 3] = blk.transpose(0, 2, 1)
        return _val(q), Jl, K
    # ---- 5-atom torsion2 family: J (and K) by differentiating the value expression
    # (source/IC/InvDisp.cpp:415-570 and :880-1068)
    order = 2 if want_K else 1
    a = _atoms_of(Jet.seeds(x, order))
    sinphi, cosphi, sign_test = _sincos_torsion2(a)
    s, c = sinphi.v, cosphi.v
    if t == "torsion2":
        use_cos = np.abs(s) > np.abs(c)  # :432, :904
        with np.errstate(divide="ignore", invalid="ignore"):   # the branch not taken may divide by 0
            Jc = cosphi.g / (-s)[:, None]
            Js_ = sinphi.g / c[:, None]
            J = np.where(use_cos[:, None], Jc, Js_)
            K = None
            if want_K:  # :916-918, :934-936
                Kc = (cosphi.h + sinphi.g[:, :, None] * Jc[:, None, :]) / (-s)[:, None, None]
                Ks = (sinphi.h - cosphi.g[:, :, None] * Js_[:, None, :]) / c[:, None, None]
                K = np.where(use_cos[:, None, None], Kc, Ks)
        phi = _wrap(_acos_signed(c, sign_test.v), d.min)
        return phi, J, K
    if t == "sintorsion2":
        return s, sinphi.g, sinphi.h
    if t == "costorsion2":
        return c, cosphi.g, cosphi.h
    # px / py: product rule with the analytic J of sin(theta) (:505-570, :982-1065)
    f = cosphi if t == "pxtorsion2" else sinphi
    st, Jst = _sintheta_J(_atoms_of(Jet.seeds(x, 1))[:3])
    Jsin = np.zeros((B, 15))
    Jsin[:, :9] = np.stack([cmp.v for j in Jst for cmp in j], axis=1)
    q = st.v * f.v
    J = Jsin * f.v[:, None] + st.v[:, None] * f.g
    K = None
    if want_K:
        Ksin = np.zeros((B, 15, 15))  # rows of atoms 0,1,2 only: d Jsintheta[i] / d r (:1008-1019)
        Ksin[:, :9, :] = np.stack([cmp.g for j in Jst for cmp in j], axis=1)
        K = (Ksin * f.v[:, None, None] + Jsin[:, :, None] * f.g[:, None, :]
             + f.g[:, :, None] * Jsin[:, None, :] + st.v[:, None, None] * f.h)  # :1022, :1064
    return q, J, K


# ----------------------------------------------------------------------------------------
# IntCoord / IntCoordSet (source/IC/IntCoord.cpp:53-94, source/IC/IntCoordSet.cpp:139-175)
# ----------------------------------------------------------------------------------------
class IntCoordSet:
    def __init__(self, fmt, text):
        self.ics = parse_intcoordset(fmt, text)
        self.intdim = len(self.ics)

    @classmethod
    def from_file(cls, fmt, path):
        with open(path) as f:
            return cls(fmt, f.read())

    def terms(self):
        out = []
        for i, ic in enumerate(self.ics):
            for c, d in ic:
                out.append((i, c, d.type, list(d.atoms), d.min))
        return out

    @staticmethod
    def _r(r):
        r = np.asarray(r, dtype=np.float64)
        return r[None] if r.ndim == 1 else r

    def q(self, r):
        """IntCoordSet::operator() (:139-145)"""
        r = self._r(r)
        q = np.empty((r.shape[0], self.intdim))
        for i, ic in enumerate(self.ics):
            acc = ic[0][0] * invdisp_value(ic[0][1], r)
            for c, d in ic[1:]:
                acc = acc + c * invdisp_value(d, r)
            q[:, i] = acc
        return q

    def qJK(self, r, want_K=True):
        """compute_IC_J_K (:161-175); with want_K False this is compute_IC_J (:147-159)"""
        r = self._r(r)
        B, cd = r.shape
        q = np.zeros((B, self.intdim))
        J = np.zeros((B, self.intdim, cd))
        K = np.zeros((B, self.intdim, cd, cd)) if want_K else None
        for i, ic in enumerate(self.ics):
            for c, d in ic:
                qi, Jl, Kl = invdisp_qJK(d, r, want_K)
                q[:, i] += c * qi
                idx = [3 * a + k for a in d.atoms for k in range(3)]
                for m, col in enumerate(idx):
                    J[:, i, col] += c * Jl[:, m]
                if want_K and idx:
                    for m, row in enumerate(idx):
                        for n_, col in enumerate(idx):
                            K[:, i, row, col] += c * Kl[:, m, n_]
        return (q, J, K) if want_K else (q, J)

    def qJ(self, r):
        return self.qJK(r, want_K=False)

    # ---- gradient / Hessian transforms (source/IC/IntCoordSet.cpp:178-266)
    def _cart2int_matrix(self, J):
        """(J J^T)^-1 J through potrf + potri like at::cholesky + at::cholesky_inverse (:182-186)"""
        from scipy.linalg import lapack
        out = np.empty_like(J)
        for b in range(J.shape[0]):
            JJT = J[b] @ J[b].T
            L, info = lapack.dpotrf(JJT, lower=1)
            assert info == 0, "J J^T is not positive definite (redundant internal coordinates?)"
            inv, info = lapack.dpotri(L, lower=1)
            assert info == 0
            inv = np.tril(inv) + np.tril(inv, -1).T
            out[b] = inv @ J[b]
        return out

    def gradient_cart2int_matrix(self, r):
        _, J = self.qJ(r)
        return self._cart2int_matrix(J)

    def gradient_cart2int(self, r, cartgrad):
        M = self.gradient_cart2int_matrix(r)
        return np.einsum("bij,bj->bi", M, np.asarray(cartgrad, dtype=np.float64).reshape(M.shape[0], -1))

    def gradient_int2cart(self, r, intgrad):
        _, J = self.qJ(r)
        return np.einsum("bij,bi->bj", J, np.asarray(intgrad, dtype=np.float64).reshape(J.shape[0], -1))

    def Hessian_int2cart(self, r, intgrad, intHess):
        """H_r = J^T H_q J + sum_k g_q[k] K[k]   (:263-264)"""
        _, J, K = self.qJK(r)
        B = J.shape[0]
        g = np.asarray(intgrad, dtype=np.float64).reshape(B, -1)
        H = np.asarray(intHess, dtype=np.float64).reshape(B, self.intdim, self.intdim)
        return np.einsum("bai,bac,bcj->bij", J, H, J) + np.einsum("bk,bkij->bij", g, K)

    def Hessian_cart2int(self, r, cartgrad, cartHess):
        """H_q = A^T H_r A - sum_k (A^T g_r)[k] A^T K[k] A,  A^T = (J J^T)^-1 J   (:236-244)"""
        _, J, K = self.qJK(r)
        B, n, cd = J.shape
        AT = self._cart2int_matrix(J)
        g = np.asarray(cartgrad, dtype=np.float64).reshape(B, cd)
        H = np.asarray(cartHess, dtype=np.float64).reshape(B, cd, cd)
        C = np.einsum("bai,bkij,bcj->bkac", AT, K, AT)
        gq = np.einsum("bij,bj->bi", AT, g)
        return np.einsum("bai,bij,bcj->bac", AT, H, AT) - np.einsum("bk,bkac->bac", gq, C)


# ----------------------------------------------------------------------------------------
# SAPSet value / Jacobian (source/polynomial/SAPSet.cpp:134-201, source/polynomial/SAP.cpp:93-175)
# ----------------------------------------------------------------------------------------
class SAPSet:
    def __init__(self, text, irreducible, dims):
        self.saps = parse_sapset(text)
        self.irreducible = int(irreducible)
        self.dims = [int(d) for d in dims]
        self.offsets = np.concatenate([[0], np.cumsum(self.dims)]).astype(int)
        self.sumdim = int(self.offsets[-1])
        self.nterms = len(self.saps)
        if self.irreducible != 0 and any(len(s) == 0 for s in self.saps):  # SAPSet.cpp:47-48
            raise ValueError("tchem::SAPSet::construct_orders_: Only the totally symmetric irreducible can have 0th order term")

    @classmethod
    def from_file(cls, path, irreducible, dims):
        with open(path) as f:
            return cls(f.read(), irreducible, dims)

    def _x(self, x):
        x = np.asarray(x, dtype=np.float64)
        x = x[None] if x.ndim == 1 else x
        if x.shape[1] != self.sumdim:  # SAPSet.cpp:135-140
            raise ValueError("x must have a same dimension as the coordinates")
        return x

    def value(self, x):
        """y[b, i] = product over coords_ in stored (descending) order, starting from 1 (SAP.cpp:93-99)"""
        x = self._x(x)
        y = np.ones((x.shape[0], self.nterms))
        for i, coords in enumerate(self.saps):
            for irr, idx in coords:
                y[:, i] = y[:, i] * x[:, self.offsets[irr] + idx]
        return y

    def jacobian(self, x):
        """Concatenated Jacobian_(xs, J): J[b, i, off(irr)+idx] = p x^(p-1) prod_{others} x_v^{p_v} (SAP.cpp:147-175).
        The reference multiplies the other factors in std::unordered_map iteration order
        (SAP.cpp:53-75); the order used here is ascending (irreducible, index), which differs from
        it by rounding only."""
        x = self._x(x)
        J = np.zeros((x.shape[0], self.nterms, self.sumdim))
        for i, coords in enumerate(self.saps):
            uniq = {}
            for c in coords:
                uniq[c] = uniq.get(c, 0) + 1
            keys = sorted(uniq)
            for u in keys:
                col = self.offsets[u[0]] + u[1]
                val = uniq[u] * np.power(x[:, col], uniq[u] - 1)
                for v in keys:
                    if v != u:
                        val = val * np.power(x[:, self.offsets[v[0]] + v[1]], uniq[v])
                J[:, i, col] = val
        return J

    def jacobian2nd(self, x):
        """Concatenated Jacobian2nd_(xs, K) (SAPSet.cpp:243-276): K[b, i] = symmetric Hessian of monomial i.
        Per monomial (SAP.cpp:226-263): diagonal p(p-1) x^(p-2) prod_{others} x_v^{p_v} (0 when p < 2),
        off-diagonal p_u p_v x_u^(p_u-1) x_v^(p_v-1) prod_{others}; distinct coordinates in ascending
        (irreducible, index) order = uniques_orders_, the other factors multiplied in that order;
        upper triangle mirrored into the lower one (SAPSet.cpp:271-274)."""
        x = self._x(x)
        K = np.zeros((x.shape[0], self.nterms, self.sumdim, self.sumdim))
        for i, coords in enumerate(self.saps):
            uniq = {}
            for c in coords:
                uniq[c] = uniq.get(c, 0) + 1
            keys = sorted(uniq)
            col = {u: self.offsets[u[0]] + u[1] for u in keys}
            for a, u in enumerate(keys):
                p = uniq[u]
                if p >= 2:
                    val = (p * (p - 1)) * np.power(x[:, col[u]], p - 2)
                    for v in keys:
                        if v != u:
                            val = val * np.power(x[:, col[v]], uniq[v])
                    K[:, i, col[u], col[u]] = val
                for v in keys[a + 1:]:
                    val = (p * uniq[v]) * np.power(x[:, col[u]], p - 1) * np.power(x[:, col[v]], uniq[v] - 1)
                    for w in keys:
                        if w != u and w != v:
                            val = val * np.power(x[:, col[w]], uniq[w])
                    K[:, i, col[u], col[v]] = val
                    K[:, i, col[v], col[u]] = val
        return K


# ----------------------------------------------------------------------------------------
# PolynomialSet (source/polynomial/PolynomialSet.cpp:56-179, source/polynomial/Polynomial.cpp:35-160): the
# non-symmetry-adapted sibling = a SAPSet with one irreducible
# ----------------------------------------------------------------------------------------
def polynomial_terms(dimension, order):
    """All terms up to `order` (PolynomialSet.cpp:60-111): orders ascending; inside an order the
    dimension-nary counter with former digit >= latter digit, i.e. coords sorted descending and
    the terms compared from the last coordinate to the first."""
    import itertools
    if dimension == 0:
        return [[]]
    terms = []
    for k in range(order + 1):
        for combo in itertools.combinations_with_replacement(range(dimension), k):
            terms.append(list(reversed(combo)))
    return terms


class PolynomialSet:
    def __init__(self, dimension, order=None, terms=None):
        self.dimension = int(dimension)
        self.terms = [sorted(t, reverse=True) for t in terms] if terms is not None else polynomial_terms(self.dimension, int(order))
        text = "".join("    ".join("1,%d" % (c + 1) for c in t) + "\n" for t in self.terms)
        self._sap = SAPSet(text, 0, [self.dimension])
        self.nterms = len(self.terms)

    def value(self, x):
        return self._sap.value(x)

    def jacobian(self, x):
        return self._sap.jacobian(x)

    def jacobian2nd(self, x):
        return self._sap.jacobian2nd(x)
