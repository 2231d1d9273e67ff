#!/usr/bin/env python
"""Per-phase clock cycles of the sparse-J transform kernels (block 0 / warp 0), C4. Needs the instrumented library
(tools/sparse_phases.sh)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_chemistry_b200 as T
from torch_chemistry_b200 import _capi

lib = _capi.lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = T.workloads.get("C4")
s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms)
s.check_factorisation = False
r = torch.from_numpy(w.geometries(B, seed=1)).cuda()
n, cd = s.intdim, s.cartdim
g = torch.Generator(device="cuda").manual_seed(3)
gq = torch.randn(B, n, dtype=torch.float64, device="cuda", generator=g)
gr = torch.randn(B, cd, dtype=torch.float64, device="cuda", generator=g)
Hq = torch.randn(B, n, n, dtype=torch.float64, device="cuda", generator=g); Hq = Hq + Hq.transpose(1, 2)
Hr = torch.randn(B, cd, cd, dtype=torch.float64, device="cuda", generator=g); Hr = Hr + Hr.transpose(1, 2)
names = {0: "factor: J rows + zero A", 1: "factor: J J^T pairs + border row", 2: "factor: potrf", 3: "factor: back substitution",
         4: "factor: trtri", 5: "factor: lauum", 6: "factor: write inverse",
         7: "   (of potrf: panels)", 8: "   (of potrf: trailing DMMA updates)",
         10: "hess i2c: wait H_q", 11: "hess i2c: U = J^T H_q", 12: "hess i2c: X = U J", 13: "hess i2c: tasks (g.K + next J rows)",
         14: "hess i2c: gather", 15: "hess i2c: store",
         20: "hess c2i: expand inverse + wait H_r", 21: "hess c2i: tasks", 22: "hess c2i: gather", 23: "hess c2i: M^T = J^T inv",
         24: "hess c2i: T2 = H_c M^T (DMMA)", 25: "hess c2i: H_q = M T2 (DMMA)", 26: "hess c2i: store", 27: "hess c2i: issue of the next prefetches"}
buf = (ctypes.c_ulonglong * 32)()
sbuf = (ctypes.c_ulonglong * 96)()
def run(label, fn, per):
    fn(); lib.tchem_debug_sparse_phases(buf, 1); lib.tchem_debug_sparse_steps(sbuf, 1)         # warm, reset
    fn(); lib.tchem_debug_sparse_phases(buf, 1); lib.tchem_debug_sparse_steps(sbuf, 1)
    tot = sum(buf[i] for i in range(32) if i not in (7, 8))
    print("%s: %d geometries through the probed thread, %.0f cycles per geometry" % (label, per, tot / per))
    for i in range(32):
        if buf[i]:
            print("   %-44s %8.0f cycles  %5.1f %%" % (names.get(i, str(i)), buf[i] / per, 100.0 * buf[i] / tot))
    if any(sbuf):
        print("   task phase, cycles per warp step (step index: static schedule order of the table): "
              + " ".join("%d:%.0f" % (i, sbuf[i] / per) for i in range(64) if sbuf[i]))
        print("   task phase, cycles per warp (whole list): " + " ".join("%.0f" % (sbuf[64 + i] / per) for i in range(16) if sbuf[64 + i]))
per_block = (B + 147) // 148
fw = int(os.environ.get("TCHEM_FACTOR_WARPS", "6"))
per_warp = (B + 148 * fw - 1) // (148 * fw)
run("Hessian_int2cart", lambda: s.Hessian_int2cart(r, gq, Hq), per_block)
run("gradient_cart2int (one warp of %d)" % fw, lambda: s.gradient_cart2int(r, gr), per_warp)
print("(Hessian_cart2int: factor phases per warp geometry, hess phases per block geometry)")
run("Hessian_cart2int", lambda: s.Hessian_cart2int(r, gr, Hr), per_block)
