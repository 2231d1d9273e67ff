"""Helper for test_gpu_parity.py::test_warp_group_kernel_variants: run in a subprocess because the
launch-shape knobs (TCHEM_GROUP_G, TCHEM_LINE) are read once per process. Compares q + J of the
20-atom workload and the aniline reference fixture with the committed goldens; prints OK."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch_chemistry_b200 as T  # noqa: E402
from conftest import assert_parity  # noqa: E402


def main():
    for name, natoms in (("workload_C3", 20), ("ref_aniline_default", 14)):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=True)
        s = T.IntCoordSet.from_text("default", str(g["definition"]), n_atoms=natoms)
        r = np.atleast_2d(g["r"])
        # ragged batch: 37 copies + the originals -> one full tile and a tail
        rr = np.concatenate([r] * (1 + 37 // len(r) + 1))[: max(37, len(r))]
        q, J = s.compute_IC_J(torch.from_numpy(rr).cuda())
        Jref = np.atleast_3d(g["J"]) if g["J"].ndim == 3 else g["J"][None]
        k = len(r)
        assert_parity(J[:k].cpu().numpy(), Jref, name + " J")
        # every copy of geometry i must give bit-identical rows
        Jh = J.cpu().numpy()
        for i in range(len(rr)):
            np.testing.assert_array_equal(Jh[i], Jh[i % k])
        qh = q.cpu().numpy()
        for i in range(len(rr)):
            np.testing.assert_array_equal(qh[i], qh[i % k])
    print("OK")


if __name__ == "__main__":
    main()
