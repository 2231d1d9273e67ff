"""Helper for test_gpu_parity.py::test_chained_cartesian_kernel_variants: run in a subprocess because TCHEM_JIT_BULK is read
once per process. Evaluates the C5 pair's chained Cartesian Jacobian on a seeded ragged batch and prints which specialised
kernel served it and a digest of the results."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch_chemistry_b200 as T  # noqa: E402


def main():
    w = T.workloads.get("C5")
    ic = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms)
    sap = T.SAPSet.from_text(w.sap_text, 0, w.sap_dims)
    r = torch.from_numpy(w.geometries(20011, seed=31)).cuda()
    y, g = sap.chained_cart(ic, r)
    torch.cuda.synchronize()
    log = T.kernel_log()
    served = [k for k in ("tchem_jit_chain_cart_bulk", "tchem_jit_chain_cart") if log.get(k, 0)]
    h = hashlib.sha256(y.cpu().numpy().tobytes() + g.cpu().numpy().tobytes()).hexdigest()
    print("%s %s" % (",".join(served), h))


if __name__ == "__main__":
    main()
