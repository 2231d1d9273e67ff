#!/usr/bin/env python
"""One launch of every serving kernel at (or near) its BASELINE size, for ncu captures:
    ncu --set full --import-source on --clock-control none -k regex:<name> -c 1 python tools/profile_all.py <which>
which: C2 | C3 | C3K | C4g | C4gc | C4H | C4Hc | C5 | C5K | C5r  (default: all, two passes - the second is the one to capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_chemistry_b200 as T

which = sys.argv[1:] or ["C2", "C3", "C3K", "C4g", "C4gc", "C4H", "C4Hc", "C5", "C5K", "C5r"]
f64 = dict(dtype=torch.float64, device="cuda")


def sets(key):
    w = T.workloads.get(key)
    s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms)
    s.check_factorisation = False
    return w, s


for rep in range(2):
    for k in which:
        if k == "C2":
            w, s = sets("C2"); r = torch.from_numpy(w.geometries(10 ** 6, seed=1)).cuda(); s.compute_IC_J(r)
        elif k == "C3":
            w, s = sets("C3"); r = torch.from_numpy(w.geometries(10 ** 5, seed=1)).cuda(); s.compute_IC_J(r)
        elif k == "C3K":
            w, s = sets("C3"); r = torch.from_numpy(w.geometries(8192, seed=1)).cuda(); s.compute_IC_J_K(r)
        elif k in ("C4g", "C4gc", "C4H", "C4Hc"):
            w, s = sets("C4"); B = 20000
            r = torch.from_numpy(w.geometries(B, seed=1)).cuda()
            n, cd = s.intdim, s.cartdim
            g = torch.Generator(device="cuda").manual_seed(3)
            if k == "C4g":
                s.gradient_int2cart(r, torch.randn(B, n, generator=g, **f64))
            elif k == "C4gc":
                s.gradient_cart2int(r, torch.randn(B, cd, generator=g, **f64))
            elif k == "C4H":
                Hq = torch.randn(B, n, n, generator=g, **f64)
                s.Hessian_int2cart(r, torch.randn(B, n, generator=g, **f64), Hq + Hq.transpose(1, 2))
            else:
                Hr = torch.randn(B, cd, cd, generator=g, **f64)
                s.Hessian_cart2int(r, torch.randn(B, cd, generator=g, **f64), Hr + Hr.transpose(1, 2))
        else:
            w, s = sets("C5")
            sap = T.SAPSet.from_text(w.sap_text, 0, w.sap_dims)
            r = torch.from_numpy(w.geometries(4 * 10 ** 6, seed=1)).cuda()
            if k == "C5":
                sap.chained(s, r)
            elif k == "C5r":
                sap.chained_cart(s, r)
            else:
                x = s(r).contiguous()
                sap.Jacobian2nd_cat(list(torch.split(x, list(w.sap_dims), dim=1)))
        torch.cuda.synchronize()
        del r
        torch.cuda.empty_cache()
print("kernels:", T.kernel_log())
