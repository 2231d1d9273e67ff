"""Where the specialised q+J kernel's time goes: the same launch with J (or q) left out - the arithmetic
and the staging stay, only the global stores disappear. usage: python tools/jit_parts.py [B]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_chemistry_b200 as T
from torch_chemistry_b200 import _capi as capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
w = T.workloads.get("C2")
s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms)
r = torch.from_numpy(w.geometries(B, seed=2)).cuda()
q = torch.empty(B, s.intdim, dtype=torch.float64, device="cuda")
J = torch.empty(B, s.intdim, s.cartdim, dtype=torch.float64, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
for name, qq, JJ in (("q+J", q, J), ("q only", q, None), ("J only", None, J)):
    ms = C.c_float(0)
    for _ in range(2):
        capi.check(capi.lib.tchem_ic_eval_timed(s._handle, p(r), B, p(qq), p(JJ), None, 0, 20, None, C.byref(ms)))
    print("%-8s %.4f ms" % (name, ms.value))
