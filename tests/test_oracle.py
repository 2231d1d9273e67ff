"""The oracle (oracle/tchem_oracle.py, a numpy restatement) pinned against the reference:
known answers of the reference's own tests, golden vectors produced by the reference itself
(tests/golden/make_golden.py), and - when oracle/_ref is built - the live reference library."""
import math

import numpy as np
import pytest

from conftest import EPS, assert_parity, assert_q_parity, c4_t32_inputs, dense_K, golden, terms_from_golden
from oracle import tchem_oracle as O

IC_CASES = ["ref_aniline_default", "ref_aniline_columbus7", "ref_aniline_advanced", "ref_normal_mode"]


@pytest.mark.parametrize("name", IC_CASES)
def test_parser_matches_reference_tables(name):
    g = golden(name)
    s = O.IntCoordSet(str(g["format"]), str(g["definition"]))
    mine, ref = s.terms(), terms_from_golden(g)
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3]
        assert abs(a[1] - b[1]) <= 4 * EPS * abs(b[1])
        assert a[4] == b[4]


@pytest.mark.parametrize("name", IC_CASES)
def test_oracle_q_J_K_against_reference_goldens(name):
    g = golden(name)
    s = O.IntCoordSet(str(g["format"]), str(g["definition"]))
    terms, r = terms_from_golden(g), g["r"]
    assert_parity(s.q(r), g["q_value"], name + " q (operator())", tol=1e-12) if name == "ref_aniline_advanced" \
        else assert_q_parity(s.q(r), g["q_value"], terms, r, name + " q (operator())")
    if "K" in g:
        q, J, K = s.qJK(r)
        assert_parity(K, g["K"], name + " K")
    else:
        q, J = s.qJ(r)
    assert_q_parity(q, g["q"], terms, r, name + " q (J path)")
    assert_parity(J, g["J"], name + " J")


def test_oracle_transforms_against_reference_goldens():
    g = golden("ref_aniline_default")
    s = O.IntCoordSet("default", str(g["definition"]))
    r = g["r"]
    assert_parity(s.gradient_int2cart(r, g["intgrad"]), g["cartgrad"], "gradient_int2cart")
    assert_parity(s.Hessian_int2cart(r, g["intgrad"], g["intHess"]), g["cartHess"], "Hessian_int2cart")
    assert_parity(s.gradient_cart2int_matrix(r), g["cart2int_matrix"], "gradient_cart2int_matrix", tol=1e-11)
    assert_parity(s.gradient_cart2int(r, g["cartgrad"]), g["intgrad_back"], "gradient_cart2int", tol=1e-11)
    assert_parity(s.Hessian_cart2int(r, g["cartgrad"], g["cartHess"]), g["intHess_back"], "Hessian_cart2int", tol=1e-11)
    # round trip of test/intcoord/main.cpp:120-149
    assert_parity(g["intgrad_back"], g["intgrad"], "reference round trip g", tol=1e-10)
    assert_parity(g["intHess_back"], g["intHess"], "reference round trip H", tol=1e-10)


def test_oracle_K_on_64_reference_geometries_C3():
    """round 2: the reference's K on 64 geometries of BASELINE config 3 (and on the normal_mode fixture)"""
    for name in ("workload_C3_K64", "ref_normal_mode_K"):
        g = golden(name)
        s = O.IntCoordSet("default", str(g["definition"]))
        q, J, K = s.qJK(g["r"])
        assert_parity(J, g["J"], name + " J (oracle)")
        assert_parity(K, dense_K(g), name + " K (oracle)")


def test_oracle_transforms_on_32_reference_geometries_C4():
    """round 2: all transforms of the reference on 32 geometries of BASELINE config 4, cond(J J^T) ~ 1e4.
    int -> cart is a plain contraction (1e-12); cart -> int inverts J J^T: the same potrf / potri algorithm in numpy
    differs from the reference's LAPACK by a few cond * eps, which is where the 1e-11 below comes from."""
    g = golden("workload_C4_T32")
    s = O.IntCoordSet("default", str(g["definition"]))
    r = g["r"]
    gq, Hq = c4_t32_inputs(g)
    assert float(g["cond_JJT"].max()) < 2e4
    assert_parity(s.gradient_int2cart(r, gq), g["cartgrad"], "C4x32 gradient_int2cart (oracle)")
    assert_parity(s.Hessian_int2cart(r[:8], gq[:8], Hq[:8]), g["cartHess"][:8], "C4x32 Hessian_int2cart (oracle)")
    assert_parity(s.gradient_cart2int(r, g["cartgrad"]), g["intgrad_back"], "C4x32 gradient_cart2int (oracle)", tol=1e-11)
    assert_parity(s.gradient_cart2int_matrix(r[:8]), g["cart2int_matrix"], "C4x32 gradient_cart2int_matrix (oracle)", tol=1e-11)
    # the reference's own round trip (test/intcoord/main.cpp:120-149) at this size
    assert_parity(g["intgrad_back"], gq, "C4x32 reference round trip g", tol=1e-9)
    assert_parity(g["intHess_back"], Hq, "C4x32 reference round trip H", tol=1e-8)


def test_columbus_golden_vector_intgeom():
    """test/intcoord/main.cpp:9-21: Columbus' own 8-decimal internal geometry (lengths in Angstrom)."""
    g, c = golden("ref_aniline_intgeom"), golden("ref_aniline_columbus7")
    intgeom = g["intgeom"].copy()
    for a, b in ((0, 7), (11, 16), (21, 23)):
        intgeom[a:b] *= 1.8897261339212517
    s = O.IntCoordSet("Columbus7", str(c["definition"]))
    q = s.q(g["geom"])[0]
    assert np.linalg.norm((q - intgeom) / intgeom) < 1e-3      # the reference itself prints 5.0e-4


def test_advanced_identities():
    """test/intcoord/main.cpp:100-113: cos(bend) = cosbending, sin(torsion) = sintorsion, ..."""
    g = golden("ref_aniline_advanced")
    s = O.IntCoordSet("whatever", str(g["definition"]))
    q = s.q(g["r"])
    err = (q[:, 0] - 1.0) + (np.cos(q[:, 1]) - q[:, 2]) + (np.cos(q[:, 3]) - q[:, 4]) + (np.cos(q[:, 5]) - q[:, 6]) \
        + (np.sin(q[:, 7]) - q[:, 8]) + (np.sin(q[:, 10]) - q[:, 11]) + (np.cos(q[:, 7]) - q[:, 9]) \
        + (np.cos(q[:, 10]) - q[:, 12]) + (np.sin(q[:, 13]) - q[:, 14]) + (np.sin(q[:, 15]) - q[:, 16]) \
        + (np.sin(q[:, 17]) - q[:, 18]) + (np.cos(q[:, 17]) - q[:, 19]) + (np.sin(q[:, 1]) * q[:, 19] - q[:, 20]) \
        + (np.sin(q[:, 1]) * q[:, 18] - q[:, 21])
    assert np.abs(err).max() < 1e-7      # limited by acos conditioning at the planar geometry


def test_oracle_J_K_finite_differences():
    """test/intcoord/main.cpp:34-53 on a small set: J, K vs central differences."""
    g = golden("workload_C2")
    s = O.IntCoordSet("default", str(g["definition"]))
    r = g["r"][:1]
    q, J, K = s.qJK(r)
    h = 1e-5
    for i in range(r.shape[1]):
        rp, rm = r.copy(), r.copy()
        rp[0, i] += h
        rm[0, i] -= h
        qp, Jp = s.qJ(rp)
        qm, Jm = s.qJ(rm)
        assert np.abs((qp - qm) / (2 * h) - J[:, :, i]).max() < 1e-8
        assert np.abs((Jp - Jm) / (2 * h) - K[:, :, :, i]).max() < 1e-8


@pytest.mark.parametrize("key", ["C1", "C2", "C3", "C4", "C5"])
def test_oracle_workloads_against_reference_goldens(key):
    from torch_chemistry_b200_workloads import get
    g = golden("workload_" + key)
    w = get(key)
    assert w.ic_text == str(g["definition"])
    np.testing.assert_array_equal(w.geometries(g["r"].shape[0]), g["r"])
    s = O.IntCoordSet("default", w.ic_text)
    terms = s.terms()
    assert_q_parity(s.q(g["r"]), g["q_value"], terms, g["r"], key + " q (operator())")
    q, J = s.qJ(g["r"])
    assert_q_parity(q, g["q"], terms, g["r"], key + " q")
    assert_parity(J, g["J"], key + " J")
    if "K" in g:
        nK = g["K"].shape[0]
        assert_parity(s.qJK(g["r"][:nK])[2], g["K"], key + " K")
    if "cartHess" in g:
        B = g["cartHess"].shape[0]
        r = g["r"][:B]
        assert_parity(s.gradient_int2cart(r, g["intgrad"]), g["cartgrad"], key + " g int2cart")
        assert_parity(s.Hessian_int2cart(r, g["intgrad"], g["intHess"]), g["cartHess"], key + " H int2cart")
        assert_parity(s.gradient_cart2int(r, g["cartgrad"]), g["intgrad_back"], key + " g cart2int", tol=1e-11)
        assert_parity(s.Hessian_cart2int(r, g["cartgrad"], g["cartHess"]), g["intHess_back"], key + " H cart2int", tol=1e-11)
    if "sap_y" in g:
        sap = O.SAPSet(str(g["sap_definition"]), 0, list(g["sap_dims"]))
        assert sap.nterms == 22
        assert_parity(sap.value(g["q"]), g["sap_y"], key + " SAP value")
        assert_parity(sap.jacobian(g["q"]), g["sap_J"], key + " SAP Jacobian")


def test_sap_known_answers():
    """test/SApolynomial/main.cpp:9-25: exact answers at xs = {(2,3),(5,7)}."""
    for name, irr, answer in (("ref_sap_1", 0, [1, 2, 3, 4, 6, 9, 25, 35, 49]), ("ref_sap_2", 1, [5, 7, 10, 14, 15, 21])):
        g = golden(name)
        s = O.SAPSet(str(g["definition"]), irr, [2, 2])
        np.testing.assert_array_equal(s.value([2.0, 3.0, 5.0, 7.0])[0], np.array(answer, dtype=float))
        np.testing.assert_array_equal(g["y"][0], np.array(answer, dtype=float))
        assert_parity(s.value(g["x"]), g["y"], name + " value")
        assert_parity(s.jacobian(g["x"]), g["J"], name + " Jacobian_")
        assert_parity(s.jacobian(g["x"]), g["J_pow"], name + " Jacobian")


def test_sap_second_order_jacobian_goldens():
    """SAPSet::Jacobian2nd_(xs, K) of the reference (goldens from oracle/_ref) vs the numpy restatement,
    plus the defining property: K is the derivative of the Jacobian (central differences)."""
    for name, xk, dk, dimk, kk in (("ref_sap_1", "x", "definition", "dims", "K2"), ("ref_sap_2", "x", "definition", "dims", "K2"),
                                   ("workload_C5", "q", "sap_definition", "sap_dims", "sap_K2")):
        g = golden(name)
        s = O.SAPSet(str(g[dk]), int(g["irreducible"]) if "irreducible" in g.files else 0, list(g[dimk]))
        x = g[xk][:len(g[kk])]
        K = s.jacobian2nd(x)
        assert_parity(K, g[kk], name + " Jacobian2nd_")
        np.testing.assert_array_equal(K, np.swapaxes(K, -1, -2))
        h = 1e-5
        for c in range(s.sumdim):
            e = np.zeros(s.sumdim)
            e[c] = h
            fd = (s.jacobian(x + e) - s.jacobian(x - e)) / (2 * h)
            assert np.abs(fd - K[..., c]).max() < 1e-6 * max(1.0, np.abs(K).max())


def test_polynomial_set_goldens():
    """PolynomialSet(dimension, order) of the reference (oracle/_ref goldens): generated term order, value
    (known answer of test/polynomial/main.cpp:7-16 is x[0] of the (2, 4) set), Jacobian_, Jacobian2nd_."""
    g = golden("ref_polynomial")
    for d, o in ((2, 4), (3, 3), (5, 2)):
        k = "d%d_o%d_" % (d, o)
        tp, cs = g[k + "term_ptr"], g[k + "coords"]
        P = O.PolynomialSet(d, o)
        assert P.terms == [list(cs[tp[i]:tp[i + 1]]) for i in range(len(tp) - 1)]
        x = g[k + "x"]
        assert_parity(P.value(x), g[k + "y"], k + "value")
        assert_parity(P.jacobian(x), g[k + "J"], k + "Jacobian_")
        assert_parity(P.jacobian2nd(x), g[k + "K"], k + "Jacobian2nd_")
    np.testing.assert_array_equal(g["d2_o4_y"][0], [1, 2, 7, 4, 14, 49, 8, 28, 98, 343, 16, 56, 196, 686, 2401])


def test_sap_constant_term_only_in_totally_symmetric():
    with pytest.raises(ValueError):
        O.SAPSet("\n1,1\n", 1, [2])


def test_oracle_against_live_reference(have_ref):
    """seeded random batches through both the restatement and the reference library."""
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import tempfile
    from oracle import ref_lib
    from torch_chemistry_b200_workloads import get
    for key, nb, nk in (("C1", 200, 20), ("C2", 100, 6), ("C3", 24, 2)):
        w = get(key)
        with tempfile.NamedTemporaryFile("w", suffix=".def", delete=False) as f:
            f.write(w.ic_text)
        ref = ref_lib.RefIntCoordSet("default", f.name)
        s = O.IntCoordSet("default", w.ic_text)
        r = w.geometries(nb, seed=1234)
        qr, Jr = ref.qJ(r)
        q, J = s.qJ(r)
        assert_q_parity(q, qr, s.terms(), r, key + " live q")
        assert_parity(J, Jr, key + " live J")
        assert_parity(s.qJK(r[:nk])[2], ref.qJK(r[:nk])[2], key + " live K")
