"""Helper for test_gpu_parity.py::test_specialised_kernel_bulk_store_variant: TCHEM_JIT_BULK is read
once per process, so each setting runs in a subprocess. Evaluates q + J through the specialised
kernel on three table shapes - even cartdim (C2H4), odd cartdim with even row count (5 atoms, 8
rows: J and q as bulk copies), even cartdim with odd row count (4 atoms, 5 rows: q rows through the
warp copy) - checks a subsample against the oracle and prints a digest of all the bytes."""
import ctypes as C
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
from torch_chemistry_b200 import _capi as capi  # noqa: E402
from oracle import tchem_oracle as O  # noqa: E402

FIVE = """    1    1.0  stretching 1 2
    2    1.0  stretching 1 3
    3    1.0  stretching 1 4
    4    1.0  stretching 1 5
    5    1.0  bending 2 1 3
         -1.0 bending 4 1 5
    6    1.0  bending 2 1 4
    7    1.0  torsion 2 1 3 4
    8    1.0  OutOfPlane 5 1 2 3
"""
FOUR = """    1    1.0  stretching 1 2
    2    1.0  stretching 2 3
         1.0  stretching 3 4
    3    1.0  bending 1 2 3
    4    1.0  cosbending 2 3 4
    5    1.0  sintorsion 1 2 3 4
"""
BASE5 = np.array([0.0, 0.0, 0.0, 1.2, 1.1, 0.9, -1.3, 1.0, -0.8, 1.1, -1.2, -1.0, -1.0, -1.1, 1.2])
BASE4 = np.array([0.0, 0.1, 0.0, 2.1, 0.0, 0.3, 2.9, 1.9, 0.0, 4.8, 2.2, 1.4])


def main():
    w = T.workloads.get("C2")
    cases = [("C2", w.ic_text, w.natoms, w.geometries(50021, seed=31))]
    rng = np.random.default_rng(7)
    cases.append(("five", FIVE, 5, BASE5 + 0.05 * rng.standard_normal((33333, 15))))
    cases.append(("four", FOUR, 4, BASE4 + 0.05 * rng.standard_normal((20011, 12))))
    h = hashlib.sha256()
    for name, text, natoms, r in cases:
        s = T.IntCoordSet.from_text("default", text, n_atoms=natoms)
        assert capi.lib.tchem_ic_uses_specialised(s._handle, len(r)) == 1, name
        q, J = s.compute_IC_J(torch.from_numpy(np.ascontiguousarray(r)).cuda())
        q, J = q.cpu().numpy(), J.cpu().numpy()
        idx = np.arange(0, len(r), 503)
        qo, Jo = O.IntCoordSet("default", text).qJ(r[idx])
        assert_parity(J[idx], Jo, name + " J vs oracle")
        assert np.abs(q[idx] - qo).max() < 1e-9, name
        # J-only and q-only calls through the C ABI write the same bytes
        rd = torch.from_numpy(np.ascontiguousarray(r)).cuda()
        q1 = torch.zeros(len(r), s.intdim, dtype=torch.float64, device="cuda")
        J1 = torch.zeros(len(r), s.intdim, s.cartdim, dtype=torch.float64, device="cuda")
        capi.check(capi.lib.tchem_ic_eval(s._handle, C.c_void_p(rd.data_ptr()), len(r), C.c_void_p(q1.data_ptr()), None, None, 0, None))
        capi.check(capi.lib.tchem_ic_eval(s._handle, C.c_void_p(rd.data_ptr()), len(r), None, C.c_void_p(J1.data_ptr()), None, 0, None))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(q1.cpu().numpy(), q)
        np.testing.assert_array_equal(J1.cpu().numpy(), J)
        h.update(q.tobytes())
        h.update(J.tobytes())
    print("DIGEST", h.hexdigest())


if __name__ == "__main__":
    main()
