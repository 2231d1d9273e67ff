#!/usr/bin/env python
"""Per-phase clock cycles of the dense per-geometry kernels (block 0), C4.

Needs a library built with -DTCHEM_DENSE_PHASES (a scratch build, never the product):
    touch torch_chemistry_b200/csrc/ic_dense.cu && python torch_chemistry_b200/build.py -DTCHEM_DENSE_PHASES
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_chemistry_b200 as T
from torch_chemistry_b200 import _capi

lib = _capi.lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = T.workloads.get("C4")
s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms)
r = torch.from_numpy(w.geometries(B, seed=1)).cuda()
n, cd = s.intdim, s.cartdim
g = torch.Generator(device="cuda").manual_seed(3)
gq = torch.randn(B, n, dtype=torch.float64, device="cuda", generator=g)
gr = torch.randn(B, cd, dtype=torch.float64, device="cuda", generator=g)
Hq = torch.randn(B, n, n, dtype=torch.float64, device="cuda", generator=g); Hq = Hq + Hq.transpose(1, 2)
Hr = torch.randn(B, cd, cd, dtype=torch.float64, device="cuda", generator=g); Hr = Hr + Hr.transpose(1, 2)
names = {0: "J (Hq load in flight)", 1: "wait Hq", 2: "T = Hq J", 3: "X = J^T T", 4: "X += g.K", 6: "store",
         8: "J", 9: "J J^T", 10: "spd inverse", 11: "M = inv J", 12: "g_q, load H_r", 13: "Hc -= g.K", 14: "two GEMMs", 15: "store",
         16: "  potrf: diagonal blocks (1 warp)", 17: "  potrf: panel rows", 18: "  potrf: trailing update", 19: "  trtri: zero W + diagonal blocks",
         20: "  trtri: L W", 21: "  trtri: -Wd (.)", 22: "  lauum: W^T W"}
buf = (ctypes.c_ulonglong * 32)()
def run(label, fn):
    fn(); lib.tchem_debug_dense_phases(buf, 1)          # warm, reset
    fn(); lib.tchem_debug_dense_phases(buf, 1)
    tot = sum(buf[i] for i in range(32) if i != 10)      # 10 = the inverse as a whole, 16-22 its parts
    per = (B + 147) // 148
    print("%s: block 0, %d geometries, %.0f cycles per geometry" % (label, per, tot / per))
    for i in range(32):
        if buf[i]:
            print("   %-28s %8.0f cycles  %5.1f %%" % (names.get(i, str(i)), buf[i] / per, 100.0 * buf[i] / tot))
run("Hessian_int2cart", lambda: s.Hessian_int2cart(r, gq, Hq))
run("gradient_cart2int", lambda: s.gradient_cart2int(r, gr))
run("Hessian_cart2int", lambda: s.Hessian_cart2int(r, gr, Hr))
