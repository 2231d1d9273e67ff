#!/usr/bin/env python
"""One launch of each dense per-geometry transform on C4 (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_chemistry_b200 as T

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
for rep in range(2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record(); s.Hessian_int2cart(r, gq, Hq)
    ev[1].record(); s.gradient_cart2int(r, gr)
    ev[2].record(); s.Hessian_cart2int(r, gr, Hr)
    ev[3].record(); s.gradient_cart2int_matrix(r)
    ev[4].record(); torch.cuda.synchronize()
    print("B=%d  H int2cart %.3f ms | g cart2int %.3f ms | H cart2int %.3f ms | cart2int matrix %.3f ms" % (
        B, *[ev[i].elapsed_time(ev[i + 1]) for i in range(4)]))
