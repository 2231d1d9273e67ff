"""Helper for test_gpu_parity.py::test_transform_kernel_variants: run in a subprocess because the kernel selection
knobs (TCHEM_DENSE_DMMA, TCHEM_FACTOR_WARPS) are read once per process. All five transforms of the aniline fixture and
the 30-atom chain against the committed reference goldens; prints OK."""
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
    want = os.environ.get("TCHEM_EXPECT_KERNEL", "")
    for name, natoms, tol in (("ref_aniline_default", 14, 1e-12), ("workload_C4", 30, 5.2e-12)):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=True)
        s = T.IntCoordSet.from_text("default", str(g["definition"]), n_atoms=natoms)
        B = g["cartHess"].shape[0]
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        r = dev(g["r"][:B])
        gq, Hq, gr, Hr = dev(g["intgrad"]), dev(g["intHess"]), dev(g["cartgrad"]), dev(g["cartHess"])
        assert_parity(s.Hessian_int2cart(r, gq, Hq).cpu().numpy(), g["cartHess"], name + " Hessian_int2cart")
        assert_parity(s.gradient_cart2int_matrix(r).cpu().numpy(), g["cart2int_matrix"], name + " cart2int matrix", tol=tol)
        assert_parity(s.gradient_cart2int(r, gr).cpu().numpy(), g["intgrad_back"], name + " gradient_cart2int", tol=tol)
        assert_parity(s.Hessian_cart2int(r, gr, Hr).cpu().numpy(), g["intHess_back"], name + " Hessian_cart2int", tol=tol)
    log = T.kernel_log()
    assert any(k.startswith(want) for k in log), (want, log)
    print("OK")


if __name__ == "__main__":
    main()
