import sys, ctypes as C, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_chemistry_b200 as T
lib = T._lib
for key, n in (("C1", 10000), ("C2", 10000), ("C1", 100000)):
    w = T.workloads.get(key)
    s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms)
    r = torch.from_numpy(w.geometries(n, seed=1)).cuda()
    q = torch.empty(n, s.intdim, dtype=torch.float64, device="cuda"); J = torch.empty(n, s.intdim, s.cartdim, dtype=torch.float64, device="cuda")
    ms = C.c_float(0)
    for _ in range(2):
        T._capi.check(lib.tchem_ic_eval_timed(s._handle, C.c_void_p(r.data_ptr()), n, C.c_void_p(q.data_ptr()), C.c_void_p(J.data_ptr()), None, 0, 500, None, C.byref(ms)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500): s.compute_IC_J(r, out=(q, J))
    torch.cuda.synchronize()
    py = (time.perf_counter() - t0) / 500 * 1e3
    print("%s B=%d: C-loop %.4f ms/launch, python-loop %.4f ms/call" % (key, n, ms.value, py))
    # the same call captured once in a CUDA graph and replayed (what a caller looping over small batches would do)
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        s.compute_IC_J(r, out=(q, J)); side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20): s.compute_IC_J(r, out=(q, J))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(25): g.replay()
    torch.cuda.synchronize()
    print("%s B=%d: CUDA graph of 20 launches: %.4f ms/launch" % (key, n, (time.perf_counter() - t0) / 500 * 1e3))
