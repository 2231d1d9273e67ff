"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's normal-mode analysis
(source/chem/normal_mode.cpp): CartNormalMode::kernel (:23-76), IntNormalMode::kernel (:109-139, the GF method),
SANormalMode::kernel (:183-222).

Pinning: the reference's chem module cannot be built here (its My_dsygv is a Fortran shim around LAPACK dsygv and
there is no Fortran compiler; at::symeig no longer exists in LibTorch 2.x - SURVEY.md section 8c), so there is NO
reference-produced golden for this path: **parity unpinned against the reference binary**. What pins it instead:
scipy.linalg.eigh(type=3) IS LAPACK dsygv(itype=3), the routine the reference calls (source/FORTRAN/*.f90), and the
reference's own acceptance test (test/normal_mode/main.cpp:68-72: Cartesian, internal and symmetry-adapted analyses give
the same frequencies on its fixture) is reproduced in tests/test_normal_mode.py on the fixture's own files."""
import numpy as np
import scipy.linalg


def signed_sqrt(x):
    """frequency^2 -> frequency, negative if imaginary (normal_mode.cpp:49-53, :118-122)"""
    return np.where(x > 0.0, np.sqrt(np.abs(x)), -np.sqrt(np.abs(x)))


def cart_normal_mode(masses, hessian):
    """CartNormalMode::kernel: mass-weighted Hessian -> eigenpairs, the 6 smallest |eigenvalue| ruled out, frequencies
    ascending; modes in rows, un-mass-weighted. Returns (frequency[3N-6], cartmode[3N-6][3N])"""
    m = np.repeat(np.asarray(masses, dtype=np.float64), 3)
    s = np.sqrt(m)
    H = np.asarray(hessian, dtype=np.float64) / s[:, None] / s[None, :]
    w, v = np.linalg.eigh(H)
    keep = np.argsort(np.abs(w), kind="stable")[6:]
    freq = signed_sqrt(w[keep])
    mode = v.T[keep] / s[None, :]
    order = np.argsort(freq, kind="stable")
    return freq[order], mode[order]


def int_normal_mode(masses, J, H):
    """IntNormalMode::kernel: G = J M^-1 J^T, dsygv type 3 (G H x = lambda x, x^T G^-1 x = 1), Linv = intmode G^-1,
    cartmode = intmode (J J^T)^-1 J. Returns (frequency[n], intmode[n][n] rows, Linv[n][n], cartmode[n][3N])"""
    J = np.asarray(J, dtype=np.float64)
    H = np.asarray(H, dtype=np.float64)
    minv = np.repeat(1.0 / np.asarray(masses, dtype=np.float64), 3)
    G = (J * minv[None, :]) @ J.T
    w, v = scipy.linalg.eigh(H, G, type=3)
    intmode = v.T
    Linv = intmode @ np.linalg.inv(G)
    cartmode = intmode @ (np.linalg.inv(J @ J.T) @ J)
    return signed_sqrt(w), intmode, Linv, cartmode


def sa_normal_mode(masses, Js, Hs):
    """SANormalMode::kernel: the same per irreducible"""
    return [int_normal_mode(masses, J, H) for J, H in zip(Js, Hs)]
