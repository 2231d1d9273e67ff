"""Helper for test_gpu_parity.py::test_gradient_int2cart_pair_kernel: TCHEM_GRAD_PAIR is read once per
process. gradient_int2cart on the 30-atom workload (odd and even batch sizes: the lone last geometry of
an odd batch is the pair kernel's edge case) and on the aniline reference fixture, checked against the
oracle / the reference goldens; prints a digest of all result bytes."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch_chemistry_b200 as T  # noqa: E402
from conftest import assert_parity  # noqa: E402
from oracle import tchem_oracle as O  # noqa: E402


def main():
    h = hashlib.sha256()
    w = T.workloads.get("C4")
    s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms)
    orc = O.IntCoordSet("default", w.ic_text)
    rng = np.random.default_rng(17)
    for n in (1, 2, 1001, 4096):
        r = w.geometries(n, seed=40 + n)
        gq = rng.standard_normal((n, s.intdim))
        gr = s.gradient_int2cart(torch.from_numpy(r).cuda(), torch.from_numpy(gq).cuda()).cpu().numpy()
        idx = np.unique(np.concatenate([np.arange(0, n, 97), [n - 1]]))
        assert_parity(gr[idx], orc.gradient_int2cart(r[idx], gq[idx]), "C4 gradient_int2cart vs oracle (n=%d)" % n)
        h.update(gr.tobytes())
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_aniline_default.npz"), allow_pickle=True)
    a = T.IntCoordSet.from_text("default", str(g["definition"]), n_atoms=14)
    k = len(g["intgrad"])
    gr = a.gradient_int2cart(torch.from_numpy(g["r"][:k]).cuda(), torch.from_numpy(g["intgrad"]).cuda()).cpu().numpy()
    assert_parity(gr, g["cartgrad"], "aniline gradient_int2cart vs the reference")
    h.update(gr.tobytes())
    print("DIGEST", h.hexdigest())


if __name__ == "__main__":
    main()
