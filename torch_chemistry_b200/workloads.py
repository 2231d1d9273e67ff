"""Synthetic workloads of BASELINE.json's configs.

Pure host-side input builders (numpy only, no device code): base geometries in bohr,
internal-coordinate definition text in the reference's default format
(source/IC/IntCoordSet.cpp:56-117 in the reference tree) and seeded perturbed batches.
Shared by bench.py, the parity tests and the golden-vector generator so that every party
evaluates the same bits.
"""
import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Workload:
    name: str
    natoms: int
    base: np.ndarray          # [3*natoms] bohr
    ic_text: str              # default-format IntCoordSet definition
    sigma: float              # std of the i.i.d. Gaussian Cartesian perturbation
    seed: int
    batch: int                # BASELINE batch size
    sap_text: str = ""        # SAP term file (C5 only)
    sap_dims: list = field(default_factory=list)

    @property
    def cartdim(self):
        return 3 * self.natoms

    def geometries(self, n=None, seed=None, dtype=np.float64):
        """[n, cartdim] perturbed geometries, reproducible from (seed, n)."""
        n = self.batch if n is None else int(n)
        rng = np.random.default_rng(self.seed if seed is None else seed)
        out = rng.standard_normal((n, self.cartdim)) * self.sigma
        out += self.base[None, :]
        return out.astype(dtype, copy=False)


def _fmt_ic(ics):
    """ics: list of ICs; each IC a list of (coeff, type, atoms(1-based), min or None)."""
    lines = []
    for i, ic in enumerate(ics):
        for k, (c, typ, atoms, mn) in enumerate(ic):
            head = "%6d" % (i + 1) if k == 0 else " " * 6
            line = head + "%12.6f%14s" % (c, typ) + "".join("%6d" % a for a in atoms)
            if mn is not None:
                line += "    %.16f" % mn
            lines.append(line)
    return "\n".join(lines) + "\n"


def h2o():
    """C1: 3 atoms, 2 stretches + 1 bend."""
    base = np.array([0.0, 0.0, 0.0, 1.4305, 1.1079, 0.0, -1.4305, 1.1079, 0.0])
    ics = [[(1.0, "stretching", (1, 2), None)],
           [(1.0, "stretching", (1, 3), None)],
           [(1.0, "bending", (2, 1, 3), None)]]
    return Workload("C1_H2O_q+J", 3, base, _fmt_ic(ics), 0.05, 1, 10 ** 4)


def c2h4():
    """C2: 6 atoms, 12 symmetry-adapted ("SAS") linear combinations, 41 invariant displacements.
    Base geometry twisted 30 deg about C=C and each CH2 wagged 10 deg so that no torsion sits at
    0 or pi (the reference's acos-based torsion is ill-conditioned there)."""
    cc, ch, hcc = 2.530, 2.054, math.radians(121.3)
    twist, wag = math.radians(30.0), math.radians(10.0)
    z = np.array([0.0, 0.0, 1.0])
    C1 = -0.5 * cc * z
    C2 = +0.5 * cc * z

    def ch2(C, toward, phi, w):
        e = np.array([math.cos(phi), math.sin(phi), 0.0])
        n = np.cross(toward, e)
        axis = math.cos(w) * toward + math.sin(w) * n
        return [C + ch * (s * math.sin(hcc) * e + math.cos(hcc) * axis) for s in (+1.0, -1.0)]

    H3, H4 = ch2(C1, +z, -0.5 * twist, wag)
    H5, H6 = ch2(C2, -z, +0.5 * twist, wag)
    base = np.concatenate([C1, C2, H3, H4, H5, H6])
    s4 = [(+1, +1, +1, +1), (+1, +1, -1, -1), (+1, -1, +1, -1), (+1, -1, -1, +1)]
    str_atoms = [(1, 3), (1, 4), (2, 5), (2, 6)]
    bend_atoms = [(3, 1, 2), (4, 1, 2), (5, 2, 1), (6, 2, 1)]
    ics = [[(1.0, "stretching", (1, 2), None)]]
    for sg in s4:
        ics.append([(float(s), "stretching", a, None) for s, a in zip(sg, str_atoms)])
    for sg in s4:
        ics.append([(float(s), "bending", a, None) for s, a in zip(sg, bend_atoms)])
    # torsion combination: pairs near 30 deg keep the default range, pairs near 210 deg use min = 0
    tors = []
    for a, d in [(3, 5), (3, 6), (4, 5), (4, 6)]:
        tors.append((a, 1, 2, d))
    ics.append([])
    for t in tors:
        phi = _dihedral(base, [i - 1 for i in t])
        ics[-1].append((1.0, "torsion", t, 0.0 if abs(phi) > 0.5 * math.pi else None))
    oop = [(2, 1, 3, 4), (1, 2, 5, 6)]
    ics.append([(1.0, "OutOfPlane", oop[0], None), (1.0, "OutOfPlane", oop[1], None)])
    ics.append([(1.0, "OutOfPlane", oop[0], None), (-1.0, "OutOfPlane", oop[1], None)])
    return Workload("C2_C2H4_q+J", 6, base, _fmt_ic(ics), 0.05, 2, 10 ** 6)


def _dihedral(r, idx):
    p = [r[3 * i:3 * i + 3] for i in idx]
    b1, b2, b3 = p[1] - p[0], p[2] - p[1], p[3] - p[2]
    n1, n2 = np.cross(b1, b2), np.cross(b2, b3)
    return math.atan2(np.dot(np.cross(n1, n2), b2 / np.linalg.norm(b2)), np.dot(n1, n2))


def _helix(natoms, bond=2.9, angle=math.radians(111.0), dihedral=math.radians(65.0)):
    """Chain built atom by atom from (bond, angle, dihedral)."""
    pts = [np.zeros(3), np.array([bond, 0.0, 0.0]),
           np.array([bond - bond * math.cos(angle), bond * math.sin(angle), 0.0])]
    while len(pts) < natoms:
        a, b, c = pts[-3], pts[-2], pts[-1]
        bc = (c - b) / np.linalg.norm(c - b)
        n = np.cross(b - a, bc); n /= np.linalg.norm(n)
        m = np.cross(n, bc)
        d2 = np.array([-bond * math.cos(angle), bond * math.sin(angle) * math.cos(dihedral),
                       bond * math.sin(angle) * math.sin(dihedral)])
        pts.append(c + d2[0] * bc + d2[1] * m + d2[2] * n)
    return np.concatenate(pts[:natoms])


def chain(natoms, name, sigma=0.03, seed=3, batch=10 ** 5, oop_variant=False):
    """C3 / C4: helical chain with Z-matrix internal coordinates:
    (N-1) stretches + (N-2) bends + (N-3) torsions = 3N-6."""
    base = _helix(natoms)
    ics = []
    for i in range(2, natoms + 1):
        ics.append([(1.0, "stretching", (i, i - 1), None)])
    for i in range(3, natoms + 1):
        ics.append([(1.0, "bending", (i, i - 1, i - 2), None)])
    for i in range(4, natoms + 1):
        if oop_variant and i % 5 == 0:
            # bond (i-1)->(i) out of the plane (i-1, i-2, i-3): centre i-1
            ics.append([(1.0, "OutOfPlane", (i, i - 1, i - 2, i - 3), None)])
        else:
            ics.append([(1.0, "torsion", (i, i - 1, i - 2, i - 3), None)])
    return Workload(name, natoms, base, _fmt_ic(ics), sigma, seed, batch)


def chain20():
    return chain(20, "C3_chain20_q+J+K", seed=3)


def chain30():
    return chain(30, "C4_chain30_grad+hess", seed=4)


def h2o_sap():
    """C5: H2O symmetric set chained into a 22-term SAPSet (irreducible 0, dims {2, 1})."""
    w = h2o()
    ics = [[(1.0, "stretching", (1, 2), None), (1.0, "stretching", (1, 3), None)],
           [(1.0, "bending", (2, 1, 3), None)],
           [(1.0, "stretching", (1, 2), None), (-1.0, "stretching", (1, 3), None)]]
    # all totally symmetric monomials up to order 4: even power of the single B2 coordinate.
    # One term per line, 1-based "irreducible,index" tokens, ordered as SAPSet requires
    # (ascending order; within an order compared from the last coordinate to the first).
    terms = []
    for order in range(0, 5):
        mono = []
        for c in range(0, order + 1, 2):
            for b in range(0, order - c + 1):
                a = order - c - b
                coords = [(2, 1)] * c + [(1, 2)] * b + [(1, 1)] * a   # descending
                mono.append(coords)
        mono.sort(key=lambda cs: list(reversed(cs)))
        terms += mono
    lines = ["    ".join("%d,%d" % t for t in cs) for cs in terms]
    sap_text = "\n".join(lines) + "\n"
    return Workload("C5_H2O_q->SAPSet", 3, w.base, _fmt_ic(ics), 0.05, 5, 10 ** 8,
                    sap_text=sap_text, sap_dims=[2, 1])


CONFIGS = {"C1": h2o, "C2": c2h4, "C3": chain20, "C4": chain30, "C5": h2o_sap}


def get(name):
    return CONFIGS[name]()
