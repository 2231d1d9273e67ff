"""Normal-mode analysis (reference source/chem/normal_mode.cpp). The reference's chem module cannot
be built here (Fortran shim around LAPACK dsygv), so the oracle (oracle/normal_mode_oracle.py: scipy's eigh(type=3) is that
same dsygv) is pinned by the reference's OWN acceptance test on its own fixture - the Cartesian, internal-coordinate and
symmetry-adapted analyses must give the same frequencies (test/normal_mode/main.cpp:68-72) - and the GPU classes are
compared with the oracle."""
import numpy as np
import pytest

from conftest import assert_parity, golden
from oracle import normal_mode_oracle as N
from oracle import tchem_oracle as O

BOHR = 1.8897261339212517


def fixture():
    g = golden("ref_normal_mode_vib")
    s = O.IntCoordSet("default", str(g["definition"]))
    r, masses = g["r"], g["masses"]
    h = g["hessian_raw"] / 4.35974417                        # test/normal_mode/main.cpp:37-48: Columbus units -> atomic units
    for a, b in ((0, 8), (16, 19), (27, 32), (39, 41)):
        h[a:b, :] /= BOHR
        h[:, a:b] /= BOHR
    q, J = s.qJ(r[None])
    return g, s, r, masses, h, J[0]


def test_oracle_reproduces_the_reference_acceptance_test():
    g, s, r, masses, h, J = fixture()
    n = J.shape[0]
    Hc = s.Hessian_int2cart(r[None], np.zeros((1, n)), h[None])[0]
    fc, mc = N.cart_normal_mode(masses, Hc)
    fi, im, Li, cm = N.int_normal_mode(masses, J, h)
    assert np.linalg.norm(fc - fi) < 1e-12                   # the reference prints this norm "close to 0"
    a, b = (int(x) for x in g["sections"])
    sa = N.sa_normal_mode(masses, [J[:a], J[a:]], [h[:a, :a], h[a:, a:]])
    assert np.linalg.norm(fc - np.sort(np.concatenate([x[0] for x in sa]))) < 1e-12
    # normalisation: intmode rows under the G^-1 metric, Linv = (intmode^T)^-1, Cartesian modes under the mass metric
    minv = np.repeat(1.0 / masses, 3)
    G = (J * minv) @ J.T
    assert np.abs(im @ np.linalg.inv(G) @ im.T - np.eye(n)).max() < 1e-10
    assert np.abs(Li @ im.T - np.eye(n)).max() < 1e-10
    assert np.abs((mc * np.repeat(masses, 3)) @ mc.T - np.eye(n)).max() < 1e-10


def _align(x, ref):
    """eigenvectors are defined up to a sign: flip each row of x to the orientation of ref"""
    sgn = np.sign(np.sum(x * ref, axis=-1, keepdims=True))
    return x * np.where(sgn == 0, 1.0, sgn)


@pytest.mark.gpu
def test_gpu_normal_modes_against_the_oracle():
    torch = pytest.importorskip("torch")
    import torch_chemistry_b200 as T
    g, s, r, masses, h, J = fixture()
    n, cd = J.shape
    # one geometry, reference call shapes (host tensors in, host tensors out)
    gs = T.IntCoordSet.from_text("default", str(g["definition"]), n_atoms=cd // 3)
    q, Jd = gs.compute_IC_J(torch.from_numpy(r).cuda())
    Hr = gs.Hessian_int2cart(torch.from_numpy(r).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda"), torch.from_numpy(h).cuda())
    cart = T.CartNormalMode(masses, Hr)
    with pytest.raises(RuntimeError):
        cart.frequency()
    cart.kernel()
    vib = T.IntNormalMode(masses, Jd, torch.from_numpy(h).cuda())
    vib.kernel()
    fi, im, Li, cm = N.int_normal_mode(masses, J, h)
    f = vib.frequency().cpu().numpy()
    assert_parity(f, fi, "normal_mode fixture: IntNormalMode frequencies", tol=1e-11)
    assert_parity(cart.frequency().cpu().numpy(), fi, "normal_mode fixture: CartNormalMode frequencies (reference acceptance test)", tol=1e-10)
    assert_parity(_align(vib.intmode().cpu().numpy(), im), im, "normal_mode fixture: intmode", tol=1e-9)
    assert_parity(_align(vib.cartmode().cpu().numpy(), cm), cm, "normal_mode fixture: cartmode", tol=1e-9)
    assert float((vib.Linv() @ vib.intmode().transpose(0, 1) - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max()) < 1e-10
    # symmetry-adapted analysis: the two irreducibles give the same spectrum (main.cpp:51-72)
    a, b = (int(x) for x in g["sections"])
    sa = T.SANormalMode(masses, [Jd[:a], Jd[a:]], [torch.from_numpy(h[:a, :a]).cuda(), torch.from_numpy(h[a:, a:]).cuda()])
    sa.kernel()
    assert sa.NIrreds() == 2
    fs = np.sort(np.concatenate([x.cpu().numpy() for x in sa.frequencies()]))
    assert_parity(fs, fi, "normal_mode fixture: SANormalMode frequencies", tol=1e-11)
    # a batch of perturbed geometries: J from the batched kernel, one eigen-solve per geometry in ONE call
    rng = np.random.default_rng(5)
    rb = r[None] + 0.01 * rng.standard_normal((9, cd))
    qb, Jb = gs.compute_IC_J(torch.from_numpy(rb).cuda())
    Hb = torch.from_numpy(np.repeat(h[None], 9, axis=0)).cuda()
    vb = T.IntNormalMode(masses, Jb, Hb)
    vb.kernel()
    fb = vb.frequency().cpu().numpy()
    assert fb.shape == (9, n) and vb.cartmode().shape == (9, n, cd)
    Jo = O.IntCoordSet("default", str(g["definition"])).qJ(rb)[1]
    for k in range(9):
        assert_parity(fb[k], N.int_normal_mode(masses, Jo[k], h)[0], "normal_mode batch: frequencies", tol=1e-11)
    # argument checks of the reference constructors
    with pytest.raises(ValueError):
        T.IntNormalMode(masses[:-1], Jd, torch.from_numpy(h).cuda())
    with pytest.raises(ValueError):
        T.IntNormalMode(masses, Jd, torch.from_numpy(h[:, :-1].copy()).cuda())
