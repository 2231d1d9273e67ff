"""Generates the golden vectors in this directory FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
It drives oracle/_ref/libtchem_ref.so - the reference's own LibTorch CPU classes compiled from
/root/reference by oracle/Makefile - on (a) the reference's own test fixtures
(test/intcoord/input/*, test/normal_mode/input/*, test/SApolynomial/input/*) and (b) the
BASELINE.json workloads C1..C5 (torch_chemistry_b200/workloads.py), and stores inputs,
definition texts and outputs as compressed .npz files. Nothing here is read at run time on
the GPU box except the .npz files.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_lib  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("workloads", os.path.join(ROOT, "torch_chemistry_b200", "workloads.py"))
workloads = importlib.util.module_from_spec(spec)
sys.modules["workloads"] = workloads
spec.loader.exec_module(workloads)

REF = "/root/reference/test"
OUT = os.path.dirname(os.path.abspath(__file__))
BOHR = 1.8897261339212517


def read_xyz(path):
    lines = open(path).read().split("\n")
    n = int(lines[0])
    return np.array([[float(x) for x in l.split()[1:4]] for l in lines[2:2 + n]]).ravel() * BOHR


def read_columbus_geom(path):
    return np.array([[float(x) for x in l.split()[2:5]] for l in open(path).read().split("\n") if l.strip()]).ravel()


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-40s %8.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024))


def ic_case(name, fmt, def_path, geoms, rng, n_pert=2, sigma=0.02, with_K=True, transforms=False):
    text = open(def_path).read()
    ref = ref_lib.RefIntCoordSet(fmt, def_path)
    r = []
    for g in geoms:
        r.append(g)
        for _ in range(n_pert):
            r.append(g + sigma * rng.standard_normal(g.shape))
    r = np.stack(r)
    out = dict(format=np.array(fmt), definition=np.array(text), r=r, q_value=ref.q(r))
    t = ref.terms()
    out.update(term_ic=t["ic"], term_coeff=t["coeff"], term_type=t["type"], term_natoms=t["natoms"],
               term_atoms=t["atoms"], term_min=t["min"])
    if with_K:
        q, J, K = ref.qJK(r)
        out.update(q=q, J=J, K=K)
    else:
        q, J = ref.qJ(r)
        out.update(q=q, J=J)
    if transforms:
        B, n, cd = r.shape[0], ref.intdim, r.shape[1]
        gq = rng.standard_normal((B, n))
        A = rng.standard_normal((B, n, n))
        Hq = 0.5 * (A + A.transpose(0, 2, 1))
        gr = ref.gradient_int2cart(r, gq)
        Hr = ref.Hessian_int2cart(r, gq, Hq)
        out.update(intgrad=gq, intHess=Hq, cartgrad=gr, cartHess=Hr,
                   cart2int_matrix=ref.gradient_cart2int_matrix(r),
                   intgrad_back=ref.gradient_cart2int(r, gr),
                   intHess_back=ref.Hessian_cart2int(r, gr, Hr))
    save(name, **out)


def main():
    rng = np.random.default_rng(20261019)
    I = REF + "/intcoord/input/"
    slow, minb1, b1rot = read_xyz(I + "slow-1.5.xyz"), read_xyz(I + "min-B1.xyz"), read_xyz(I + "B1rot-2.2.xyz")
    geom = read_columbus_geom(I + "geom")
    intgeom = np.array([float(x) for x in open(I + "intgeom").read().split()])
    # reference fixtures (test/intcoord/main.cpp)
    ic_case("ref_aniline_default", "default", I + "IntCoordDef", [slow, minb1, b1rot], rng, n_pert=1, transforms=True)
    ic_case("ref_aniline_columbus7", "Columbus7", I + "intcfl", [geom, slow], rng, n_pert=0)
    ic_case("ref_aniline_advanced", "whatever", I + "advanced_IntCoordDef", [slow, minb1, b1rot], rng, n_pert=1)
    save("ref_aniline_intgeom", intgeom=intgeom, geom=geom)
    # test/normal_mode fixture: 48 ICs with un-normalised coefficients
    N = REF + "/normal_mode/input/"
    ic_case("ref_normal_mode", "default", N + "IntCoordDef", [read_columbus_geom(N + "geom")], rng, n_pert=1, with_K=False)
    # SAP known answers (test/SApolynomial/main.cpp:9-25)
    S = REF + "/SApolynomial/input/"
    x = np.concatenate([np.array([[2.0, 3.0, 5.0, 7.0]]), rng.standard_normal((7, 4))])
    for f, irr in (("1.in", 0), ("2.in", 1)):
        sap = ref_lib.RefSAPSet(S + f, irr, [2, 2])
        save("ref_sap_" + f.replace(".in", ""), definition=np.array(open(S + f).read()), irreducible=np.array(irr),
             dims=np.array([2, 2]), x=x, y=sap.value(x), J=sap.jacobian(x), J_pow=sap.jacobian(x, True),
             K2=sap.jacobian2nd(x))
    # BASELINE workloads
    for key, nK, nT in (("C1", 64, 16), ("C2", 48, 8), ("C3", 6, 0), ("C4", 0, 3), ("C5", 32, 0)):
        w = workloads.get(key)
        with tempfile.NamedTemporaryFile("w", suffix=".def", delete=False) as f:
            f.write(w.ic_text)
            path = f.name
        ref = ref_lib.RefIntCoordSet("default", path)
        nJ = 256 if key != "C4" else 32
        r = w.geometries(nJ)
        out = dict(definition=np.array(w.ic_text), base=w.base, r=r, q_value=ref.q(r))
        out["q"], out["J"] = ref.qJ(r)
        if nK:
            _, _, out["K"] = ref.qJK(r[:nK])
        if nT:
            B, n = nT, ref.intdim
            gq = rng.standard_normal((B, n))
            A = rng.standard_normal((B, n, n))
            Hq = 0.5 * (A + A.transpose(0, 2, 1))
            gr = ref.gradient_int2cart(r[:B], gq)
            Hr = ref.Hessian_int2cart(r[:B], gq, Hq)
            out.update(intgrad=gq, intHess=Hq, cartgrad=gr, cartHess=Hr,
                       cart2int_matrix=ref.gradient_cart2int_matrix(r[:B]),
                       intgrad_back=ref.gradient_cart2int(r[:B], gr),
                       intHess_back=ref.Hessian_cart2int(r[:B], gr, Hr))
        if w.sap_text:
            with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f:
                f.write(w.sap_text)
                spath = f.name
            sap = ref_lib.RefSAPSet(spath, 0, w.sap_dims)
            out.update(sap_definition=np.array(w.sap_text), sap_dims=np.array(w.sap_dims),
                       sap_y=sap.value(out["q"]), sap_J=sap.jacobian(out["q"]), sap_K2=sap.jacobian2nd(out["q"][:64]))
        save("workload_" + key, **out)


def add_sap_second_order():
    """`python make_golden.py sapK2`: adds SAPSet::Jacobian2nd_(xs, K) of the reference to the SAP
    goldens that already exist (every other array is kept byte for byte)."""
    for name, xkey, dkey, dimkey, irrkey, kkey, n in (("ref_sap_1", "x", "definition", "dims", "irreducible", "K2", None),
                                                      ("ref_sap_2", "x", "definition", "dims", "irreducible", "K2", None),
                                                      ("workload_C5", "q", "sap_definition", "sap_dims", None, "sap_K2", 64)):
        g = dict(np.load(os.path.join(OUT, name + ".npz"), allow_pickle=True))
        with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f:
            f.write(str(g[dkey]))
            spath = f.name
        sap = ref_lib.RefSAPSet(spath, int(g[irrkey]) if irrkey else 0, [int(d) for d in g[dimkey]])
        g[kkey] = sap.jacobian2nd(g[xkey][:n] if n else g[xkey])
        save(name, **g)


def polynomial_goldens():
    """`python make_golden.py poly`: PolynomialSet(dimension, order) of the reference - generated terms,
    value, Jacobian_, Jacobian2nd_ (known answer of test/polynomial/main.cpp:7-16 included as x[0] of (2, 4))."""
    rng = np.random.default_rng(77)
    out = {}
    for d, o in ((2, 4), (3, 3), (5, 2)):
        p = ref_lib.RefPolynomialSet(d, o)
        terms = p.terms()
        x = rng.standard_normal((8, d)) * 0.6 + 1.0
        if (d, o) == (2, 4):
            x[0] = [2.0, 7.0]
        y, J, K = p.eval(x)
        key = "d%d_o%d_" % (d, o)
        out[key + "term_ptr"] = np.cumsum([0] + [len(t) for t in terms]).astype(np.int32)
        out[key + "coords"] = np.array([c for t in terms for c in t], dtype=np.int32)
        out.update({key + "x": x, key + "y": y, key + "J": J, key + "K": K})
    save("ref_polynomial", **out)


def sparse_K(K):
    """K[B][n][cd][cd] of the reference -> (flat indices of the elements that are nonzero in ANY geometry, values).
    Everything else is exactly 0.0 in the reference's output (asserted), so the dense tensor is recoverable bit for bit."""
    B = K.shape[0]
    flat = K.reshape(B, -1)
    idx = np.flatnonzero(np.any(flat != 0.0, axis=0))
    rest = flat.copy()
    rest[:, idx] = 0.0
    assert not rest.any()
    return idx.astype(np.int64), np.ascontiguousarray(flat[:, idx])


def round2_goldens():
    """`python make_golden.py round2`: the larger reference-held samples VERDICT r01 asked for, as separate files
    (the round-1 files stay byte for byte):
      workload_C3_K64  - q, J, K of the reference on 64 geometries of C3 (K stored as pattern + values)
      workload_C4_T32  - all five transforms on 32 geometries of C4, inputs exactly representable (k / 16), with
                         cond(J J^T) per geometry
      ref_normal_mode_K - K on the reference's test/normal_mode fixture (un-normalised coefficients)"""
    rng = np.random.default_rng(20261020)
    # ---- C3: K on 64 geometries
    w = workloads.get("C3")
    with tempfile.NamedTemporaryFile("w", suffix=".def", delete=False) as f:
        f.write(w.ic_text)
        path = f.name
    ref = ref_lib.RefIntCoordSet("default", path)
    r = w.geometries(64, seed=3003)
    q, J, K = ref.qJK(r)
    kidx, kval = sparse_K(K)
    save("workload_C3_K64", definition=np.array(w.ic_text), r=r, q=q, J=J, K_idx=kidx, K_val=kval, K_shape=np.array(K.shape))
    # ---- C4: transforms on 32 geometries
    w = workloads.get("C4")
    with tempfile.NamedTemporaryFile("w", suffix=".def", delete=False) as f:
        f.write(w.ic_text)
        path = f.name
    ref = ref_lib.RefIntCoordSet("default", path)
    B, n, cd = 32, ref.intdim, w.cartdim
    r = w.geometries(B, seed=4004)
    gq16 = rng.integers(-48, 49, size=(B, n)).astype(np.int8)              # exact in float64 after / 16
    A16 = rng.integers(-40, 41, size=(B, n, n)).astype(np.int16)
    Hq16 = (A16 + A16.transpose(0, 2, 1)).astype(np.int16)                # symmetric, / 32
    gq, Hq = gq16 / 16.0, Hq16 / 32.0
    gr = ref.gradient_int2cart(r, gq)
    Hr = ref.Hessian_int2cart(r, gq, Hq)
    _, Jm = ref.qJ(r)
    cond = np.array([np.linalg.cond(Jm[b] @ Jm[b].T) for b in range(B)])
    save("workload_C4_T32", definition=np.array(w.ic_text), r=r, intgrad16=gq16, intHess32=Hq16, cartgrad=gr, cartHess=Hr,
         cart2int_matrix=ref.gradient_cart2int_matrix(r[:8]), intgrad_back=ref.gradient_cart2int(r, gr),
         intHess_back=ref.Hessian_cart2int(r, gr, Hr), cond_JJT=cond)
    # ---- normal_mode fixture with K
    N = REF + "/normal_mode/input/"
    g0 = read_columbus_geom(N + "geom")
    r = np.stack([g0, g0 + 0.02 * rng.standard_normal(g0.shape)])
    ref = ref_lib.RefIntCoordSet("default", N + "IntCoordDef")
    q, J, K = ref.qJK(r)
    kidx, kval = sparse_K(K)
    save("ref_normal_mode_K", format=np.array("default"), definition=np.array(open(N + "IntCoordDef").read()), r=r, q=q, J=J,
         K_idx=kidx, K_val=kval, K_shape=np.array(K.shape))


def normal_mode_fixture():
    """`python make_golden.py vib`: the INPUT files of the reference's test/normal_mode (definition, geometry, masses, the
    Columbus internal-coordinate Hessian as read by test/normal_mode/main.cpp:22-36) - no outputs: the reference's chem module
    cannot be built here (no Fortran compiler for My_dsygv), see oracle/normal_mode_oracle.py."""
    d = REF + "/normal_mode/input/"
    geom = [l.split() for l in open(d + "geom") if l.strip()]
    r = np.array([[float(x) for x in g[2:5]] for g in geom]).ravel()
    masses = np.array([float(g[5]) for g in geom])
    h = np.array(open(d + "hessian").read().split(), dtype=float).reshape(48, 48)
    save("ref_normal_mode_vib", definition=np.array(open(d + "IntCoordDef").read()), r=r, masses=masses, hessian_raw=h,
         sections=np.array([27, 21]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vib":
        normal_mode_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "round2":
        round2_goldens()
    elif len(sys.argv) > 1 and sys.argv[1] == "poly":
        polynomial_goldens()
    elif len(sys.argv) > 1 and sys.argv[1] == "sapK2":
        add_sap_second_order()
    else:
        main()
