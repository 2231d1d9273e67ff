#!/usr/bin/env python
"""bench.py - geometries/s of the internal-coordinate hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--extras all|none|C3,C4H,...]
                    [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic geometries. The headline (`value`,
`roofline`, `e2e`, `cpu_baseline` at the top level of the JSON line) is BASELINE.json configs[1]: C2H4,
12 SAS internal coordinates, q + Wilson-B on 10^6 geometries per GPU. `extra` holds one record per
remaining BASELINE config, measured the same way in the same run:
    C1   H2O q+B, 10^4                     C3   20 atoms q+B, 10^5        C3K  20 atoms q+B+K, 10^5 (chunks of 8192)
    C4g  30 atoms gradient int->cart, 10^5 C4gc gradient cart->int, 10^5  C4H  Hessian int->cart, 10^5
    C4Hc Hessian cart->int, 10^5           C5   r->q->SAPSet value+Jacobian, 1.25*10^7 per GPU (10^8 over 8)
    C5K  SAPSet second-order Jacobian, 1.25*10^7
    C5r  r->q->SAPSet value + CARTESIAN Jacobian (dSAP/dq) J, 1.25*10^7
    C5r2 the same plus the Cartesian second derivatives J^T (d2SAP/dq2) J + sum_k (dSAP/dq_k) K[k], 2*10^5 (composed path)
With N > 1 every rank owns its own shard of every workload (weak scaling, no collective in the data path) and
the extras default to C5 (BASELINE config 5) and C4H. One JSON line is printed by rank 0; DESIGN.md section 5.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic bytes per geometry: inputs read once + API-mandated dense outputs written once
# (SURVEY.md section 8d / DESIGN.md)
WORKLOADS = {
    # key: (what, bytes per geometry as a function of (cartdim, intdim, nsap))
    "C1": ("q+J", lambda cd, n, ns: 8 * (cd + n + n * cd)),
    "C2": ("q+J", lambda cd, n, ns: 8 * (cd + n + n * cd)),
    "C3": ("q+J", lambda cd, n, ns: 8 * (cd + n + n * cd)),
    "C3K": ("q+J+K", lambda cd, n, ns: 8 * (cd + n + n * cd + n * cd * cd)),
    "C4J": ("q+J", lambda cd, n, ns: 8 * (cd + n + n * cd)),
    "C4g": ("gradient_int2cart", lambda cd, n, ns: 8 * (cd + n + cd)),
    "C4gc": ("gradient_cart2int", lambda cd, n, ns: 8 * (cd + cd + n)),
    "C4H": ("Hessian_int2cart", lambda cd, n, ns: 8 * (cd + n + n * n + cd * cd)),
    "C4Hc": ("Hessian_cart2int", lambda cd, n, ns: 8 * (cd + cd + cd * cd + n * n)),
    "C5": ("q->SAPSet value+Jacobian", lambda cd, n, ns: 8 * (cd + ns + ns * n)),
    "C5r": ("q->SAPSet value+Cartesian Jacobian", lambda cd, n, ns: 8 * (cd + ns + ns * cd)),
    "C5r2": ("q->SAPSet value+Cartesian Jacobian+Cartesian second derivatives", lambda cd, n, ns: 8 * (cd + ns + ns * cd + ns * cd * cd)),
    "C5x": ("SAPSet value+Jacobian on given coordinates", lambda cd, n, ns: 8 * (n + ns + ns * n)),
    "C5K": ("SAPSet Jacobian2nd on given coordinates", lambda cd, n, ns: 8 * (n + ns * n * n)),
}
WL_KEY = {"C3K": "C3", "C5x": "C5", "C5K": "C5", "C5r": "C5", "C5r2": "C5", "C4J": "C4", "C4g": "C4", "C4gc": "C4", "C4H": "C4", "C4Hc": "C4"}
DEFAULT_BATCH = {"C1": 10 ** 4, "C2": 10 ** 6, "C3": 10 ** 5, "C3K": 10 ** 5, "C4J": 10 ** 5, "C4g": 10 ** 5, "C4gc": 10 ** 5,
                 "C4H": 10 ** 5, "C4Hc": 10 ** 5, "C5": 12500000, "C5x": 12500000, "C5K": 12500000, "C5r": 12500000, "C5r2": 200000}
K_CHUNK = 8192          # q+J+K: 1.55 MB of K per 20-atom geometry -> the batch is walked in chunks
EXTRAS_1GPU = ["C1", "C3", "C3K", "C4g", "C4gc", "C4H", "C4Hc", "C5", "C5K", "C5r", "C5r2"]
EXTRAS_NGPU = ["C5", "C4H"]
# Dense-equivalent flops per geometry of the contraction-bound transforms: what the reference's algorithm costs
# with dense operands, counted once, LAPACK conventions (potrf n^3/3 + potri 2n^3/3 = n^3).
#   C4H : H_q J (2 n n cd) + J^T (.) (2 cd n cd) + sum_k g_k K_k (2 n cd cd)
#   C4gc: J J^T (2 n n cd) + potrf/potri (n^3) + J g_r (2 n cd) + inv (.) (2 n n)        [no M = inv J product is needed]
#   C4Hc: J J^T + potrf/potri + M = inv J (2 n n cd) + sum_k g_k K_k (2 n cd cd) + M Hc M^T (2 cd cd n + 2 n cd n)
# The kernels exploit J's structural sparsity and execute far fewer flops than this; `flops_executed_per_geometry`
# (counted from the definition table) is reported next to it.
FLOPS = {
    "C4H": lambda cd, n: 2 * n * n * cd + 2 * cd * n * cd + 2 * n * cd * cd,
    "C4gc": lambda cd, n: 2 * n * n * cd + n ** 3 + 2 * n * cd + 2 * n * n,
    "C4Hc": lambda cd, n: 2 * n * n * cd + n ** 3 + 2 * n * n * cd + 2 * n * cd * cd + 2 * cd * cd * n + 2 * n * cd * n,
}

_FP64_PEAK = {}


def measured_fp64_peak(device):
    """cuBLAS DGEMM 8192^3 through torch.matmul (TFLOP/s): best of 5 (burst) and back to back for ~1 s (sustained);
    same method as MEASURED_PEAKS.json's bf16 number. Measured once per process."""
    if _FP64_PEAK:
        return _FP64_PEAK
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    reps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    _FP64_PEAK.update(burst=best, sustained=reps * 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return _FP64_PEAK


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML, 50 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def load_workloads_module():
    import importlib.util
    if "_wl" in sys.modules:
        return sys.modules["_wl"]
    spec = importlib.util.spec_from_file_location("_wl", os.path.join(ROOT, "torch_chemistry_b200", "workloads.py"))
    wl = importlib.util.module_from_spec(spec)
    sys.modules["_wl"] = wl
    spec.loader.exec_module(wl)
    return wl


# ------------------------------------------------------------------------------------------
# CPU baseline: the reference's own LibTorch implementation (oracle/_ref), one process per core.
# One pool serves every workload of the run: a worker builds the reference objects of a workload the
# first time it sees it (table construction is not timed).
# ------------------------------------------------------------------------------------------
_WORKER = {}


def _cpu_objects(kind, wkey, what):
    key = (wkey, "SAP" in what)
    if key in _WORKER:
        return _WORKER[key]
    w = load_workloads_module().get(wkey)
    if kind == "reference":
        from oracle import ref_lib
        with tempfile.NamedTemporaryFile("w", suffix=".def", delete=False) as f:
            f.write(w.ic_text)
        s = ref_lib.RefIntCoordSet("default", f.name)
        sap = None
        if "SAP" in what:
            with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f2:
                f2.write(w.sap_text)
            sap = ref_lib.RefSAPSet(f2.name, 0, w.sap_dims)
    else:
        from oracle import tchem_oracle as O
        s = O.IntCoordSet("default", w.ic_text)
        sap = O.SAPSet(w.sap_text, 0, w.sap_dims) if "SAP" in what else None
    _WORKER[key] = (s, sap)
    return s, sap


def _cpu_worker(job):
    import time as _t
    kind, wkey, what, r = job
    s, sap = _cpu_objects(kind, wkey, what)
    t0 = _t.perf_counter()
    if what == "q+J":
        s.qJ(r)
    elif what == "q+J+K":
        s.qJK(r)
    elif what == "gradient_int2cart":
        s.gradient_int2cart(r, np.ones((r.shape[0], s.intdim)))
    elif what == "gradient_cart2int":
        s.gradient_cart2int(r, np.ones((r.shape[0], r.shape[1])))
    elif what == "Hessian_int2cart":
        s.Hessian_int2cart(r, np.ones((r.shape[0], s.intdim)), np.tile(np.eye(s.intdim), (r.shape[0], 1, 1)))
    elif what == "Hessian_cart2int":
        s.Hessian_cart2int(r, np.ones((r.shape[0], r.shape[1])), np.tile(np.eye(r.shape[1]), (r.shape[0], 1, 1)))
    elif "second derivatives" in what:
        q, J, K = s.qJK(r)                     # the same composition, second order
        sap.value(q)
        g = sap.jacobian(q)
        np.einsum("btk,bkc->btc", g, J)
        np.einsum("bka,btkl,blc->btac", J, sap.jacobian2nd(q), J) + np.einsum("btk,bkac->btac", g, K)
    elif "Cartesian" in what:
        q, J = s.qJ(r)                         # what callers of the reference do by hand (test/normal_mode/main.cpp:56-61)
        sap.value(q)
        np.einsum("btk,bkc->btc", sap.jacobian(q), J)
    elif what.startswith("SAPSet"):
        q = s.qJ(r)[0]
        t0 = _t.perf_counter()                 # the coordinates are the caller's: only the SAPSet call is timed
        if "Jacobian2nd" in what:
            sap.jacobian2nd(q)
        else:
            sap.value(q)
            sap.jacobian(q)
    else:
        q = s.qJ(r)[0]
        sap.value(q)
        sap.jacobian(q)
    return _t.perf_counter() - t0


class CpuBaseline:
    """The reference CPU path (oracle/_ref: the reference's own LibTorch classes; falls back to the
    numpy port when that library is absent) on one single-threaded process per host core."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        from oracle import ref_lib
        self.kind = "reference" if ref_lib.available() else "port"
        try:
            avail = len(os.sched_getaffinity(0))
        except Exception:
            avail = os.cpu_count() or 1
        self.cores = 1 if self.kind == "port" else min(cores or avail, 128)
        self.pool = mp.get_context("spawn").Pool(self.cores)

    def run(self, wkey, what, per_core, seed=4242):
        w = load_workloads_module().get(wkey)
        warm = w.geometries(2, seed=1)
        self.pool.map(_cpu_worker, [(self.kind, wkey, what, warm)] * self.cores, chunksize=1)     # tables, first-call allocations
        r = w.geometries(per_core * self.cores, seed=seed)
        jobs = [(self.kind, wkey, what, r[i * per_core:(i + 1) * per_core]) for i in range(self.cores)]
        t0 = time.perf_counter()
        times = self.pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        # SAPSet on given coordinates: the workers time the SAPSet call alone (the coordinates are the caller's)
        span = max(times) if what.startswith("SAPSet") else wall
        return {"value": per_core * self.cores / span, "unit": "geometries/s", "cores": self.cores, "kind": self.kind,
                "sample": "%d geometries (%d per process, %d single-threaded processes), %s; wall %.2f s, slowest process %.2f s"
                          % (per_core * self.cores, per_core, self.cores, what, wall, max(times))}

    def close(self):
        self.pool.close()
        self.pool.join()


def per_core_sample(key, headline):
    # headline: roughly 10-20 s of CPU work per process at the reference's measured per-geometry cost;
    # extras: a few seconds each, so that the default run stays within minutes
    full = {"C1": 20000, "C2": 4000, "C3": 1500, "C3K": 24, "C4J": 1000, "C4g": 1000, "C4gc": 1000, "C4H": 16, "C4Hc": 16,
            "C5": 20000, "C5x": 20000, "C5K": 20000, "C5r": 20000, "C5r2": 2000}[key]
    return full if headline else max(4, full // 4)


# ------------------------------------------------------------------------------------------
def run_reference_arm(args, metric, config):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    what = WORKLOADS[args.workload][0]
    wkey = WL_KEY.get(args.workload, args.workload)
    per_core = max(4, per_core_sample(args.workload, True) // 8)      # ~1-2 s of CPU work per step
    vals = []
    base = None
    cb = CpuBaseline()
    for i in range(args.warmup + args.steps):
        base = cb.run(wkey, what, per_core, seed=100 + i)
        if i >= args.warmup:
            vals.append(per_core * base["cores"] / base["value"])
    cb.close()
    v = per_core * base["cores"] / float(np.mean(vals))
    base["value"] = v
    line = {"metric": metric, "value": v, "unit": "geometries/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * per_core * base["cores"] / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "impl": "reference", "cpu_baseline": base,
            "e2e": {"value": v, "unit": "geometries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


class Bench:
    """One process = one GPU. measure(key) times one workload and returns its record."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import torch_chemistry_b200 as T
        self.torch, self.dist, self.T, self.args = torch, dist, T, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.device = torch.device("cuda", self.local)
        self.f64 = dict(dtype=torch.float64, device=self.device)
        self.flush_buf = None
        self.cpu = None
        self.pcie = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def pcie_ceiling(self):
        """plain pinned cudaMemcpyAsync ceilings of this GPU's host link (GB/s), all ranks copying at once:
        what an end-to-end number can reach at best"""
        if self.pcie is not None:
            return self.pcie
        torch = self.torch
        nbytes = 1 << 30
        h = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
        d = torch.empty(nbytes // 8, **self.f64)
        out = {}
        for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
            dst.copy_(src, non_blocking=True)
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt = self.max_over_ranks(time.perf_counter() - t0)
            out[name + "_GBps_per_gpu"] = 3 * nbytes / dt / 1e9
        out["how"] = "3 x 1 GiB pinned cudaMemcpyAsync per direction, all %d rank(s) at once, max over ranks" % self.world
        self.pcie = out
        return out

    # ------------------------------------------------------------------------------------------
    def measure(self, key, B, steps, warmup, headline):
        torch, T, device, f64 = self.torch, self.T, self.device, self.f64
        args = self.args
        wl = load_workloads_module()
        wkey = WL_KEY.get(key, key)
        w = wl.get(wkey)
        what, bytes_fn = WORKLOADS[key]
        ic = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms, device=self.local)
        ic.check_factorisation = False          # no synchronisation inside the timed region (polled after it)
        sap = T.SAPSet.from_text(w.sap_text, 0, w.sap_dims, device=self.local) if w.sap_text else None
        n, cd = ic.intdim, ic.cartdim
        ns = sap.nterms if sap else 0
        bytes_per_geom = bytes_fn(cd, n, ns)

        # synthetic shard of this rank: same generator, rank-dependent seed (shard-invariant definition)
        r_host = torch.from_numpy(w.geometries(B, seed=w.seed + 1000 * self.rank)).pin_memory()
        r_dev = r_host.to(device)
        gen = torch.Generator(device=device).manual_seed(77 + self.rank)
        rand = lambda *s: torch.randn(*s, generator=gen, **f64)
        gq_dev = rand(B, n) if what in ("gradient_int2cart", "Hessian_int2cart") else None
        gr_dev = rand(B, cd) if what in ("gradient_cart2int", "Hessian_cart2int") else None
        Hq_dev = Hr_dev = None
        if what == "Hessian_int2cart":
            Hq_dev = rand(B, n, n)
            Hq_dev = 0.5 * (Hq_dev + Hq_dev.transpose(1, 2)).contiguous()
        if what == "Hessian_cart2int":
            Hr_dev = rand(B, cd, cd)
            Hr_dev = 0.5 * (Hr_dev + Hr_dev.transpose(1, 2)).contiguous()

        launches_per_step = 1
        if what == "q+J":
            outs = [torch.empty(B, n, **f64), torch.empty(B, n, cd, **f64)]
            step = lambda: ic.compute_IC_J(r_dev, out=outs)
        elif what == "q+J+K":
            # the whole batch in chunks of K_CHUNK geometries through one chunk-sized result buffer
            nb = min(B, K_CHUNK)
            outs = [torch.empty(nb, n, **f64), torch.empty(nb, n, cd, **f64), torch.empty(nb, n, cd, cd, **f64)]
            cuts = [(b0, min(B, b0 + nb)) for b0 in range(0, B, nb)]
            launches_per_step = len(cuts)

            def step():
                for b0, b1 in cuts:
                    ic.compute_IC_J_K(r_dev[b0:b1], out=[o[:b1 - b0] for o in outs])
        elif what == "gradient_int2cart":
            step = lambda: ic.gradient_int2cart(r_dev, gq_dev)
        elif what == "gradient_cart2int":
            step = lambda: ic.gradient_cart2int(r_dev, gr_dev)
        elif what == "Hessian_int2cart":
            out_H = torch.empty(B, cd, cd, **f64)
            step = lambda: ic.Hessian_int2cart(r_dev, gq_dev, Hq_dev, out=out_H)
        elif what == "Hessian_cart2int":
            out_H = torch.empty(B, n, n, **f64)
            step = lambda: ic.Hessian_cart2int(r_dev, gr_dev, Hr_dev, out=out_H)
        elif key == "C5K":
            import ctypes as C
            x_dev = ic(r_dev).contiguous()
            K2 = torch.empty(B, ns, n, n, **f64)
            step = lambda: T._capi.check(T._lib.tchem_sap_jacobian2nd(sap._handle, C.c_void_p(x_dev.data_ptr()), B, C.c_void_p(K2.data_ptr()),
                                                                      C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        elif key == "C5r":
            step = lambda: sap.chained_cart(ic, r_dev)
        elif key == "C5r2":
            step = lambda: sap.chained_cart(ic, r_dev, second_order=True)
        elif key == "C5x":
            x_dev = ic(r_dev).contiguous()                      # the coordinates a caller would bring
            souts = sap.eval_cat(x_dev)
            step = lambda: sap.eval_cat(x_dev, out=souts)
        else:
            couts = sap.chained(ic, r_dev)
            step = lambda: sap.chained(ic, r_dev, out=list(couts))
        need_flush = B * bytes_per_geom <= 126e6
        if need_flush and self.flush_buf is None:
            self.flush_buf = torch.empty(160 * 1024 * 1024 // 8, **f64)

        jit0 = T.jit_stats()
        t_first = time.perf_counter()
        step()
        torch.cuda.synchronize()
        first_call_s = time.perf_counter() - t_first
        jit1 = T.jit_stats()
        for _ in range(max(warmup - 1, 0)):
            step()
        self.barrier()
        sampler = ClockSampler(self.local)
        sampler.start()
        log0 = T.kernel_log()
        launches0 = T.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            if need_flush:
                self.flush_buf.zero_()
            ev[i][0].record()
            step()
            ev[i][1].record()
        self.barrier()
        launches = T.launch_count() - launches0
        log1 = T.kernel_log()
        sampler.stop_flag = True
        sampler.join()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms_max = self.max_over_ranks(float(sum(step_ms)))
        value = self.world * B * steps / (total_ms_max * 1e-3)
        kernels = {k: v - log0.get(k, 0) for k, v in log1.items() if v - log0.get(k, 0) > 0}
        factor_failures = ic.factor_status() if "cart2int" in what else 0

        # ---- e2e: host buffers through the public API, copies inside the timed region --------------
        e2e = None
        if not args.no_e2e and key not in ("C5x", "C5K", "C5r", "C5r2"):
            pin = lambda *s: torch.empty(s, dtype=torch.float64).pin_memory()
            Be = B
            if sap is not None:
                houts = [pin(B, ns), pin(B, ns, n)]
                ins = [r_host]
                call = lambda: sap.chained(ic, r_host, out=houts)
                api = "SAPSet.chained(IntCoordSet, r) with pinned host tensors -> tchem_ic_sap_eval_host"
            elif what == "q+J":
                houts = [pin(B, n), pin(B, n, cd)]
                ins = [r_host]
                call = lambda: ic.compute_IC_J(r_host, out=houts)
                api = "IntCoordSet.compute_IC_J(r) with pinned host tensors -> tchem_ic_eval_host"
            elif what == "q+J+K":
                Be = min(B, K_CHUNK)            # one chunk: 12.7 GB of results per call
                houts = [pin(Be, n), pin(Be, n, cd), pin(Be, n, cd, cd)]
                ins = [r_host[:Be]]
                call = lambda: ic.compute_IC_J_K(r_host[:Be], out=houts)
                api = "IntCoordSet.compute_IC_J_K(r) with pinned host tensors -> tchem_ic_eval_host (one %d-geometry chunk)" % Be
            else:
                Be = min(B, 20000) if what.startswith("Hessian") else B      # 1.3 GB of operands / results per direction
                src = {"gradient_int2cart": [gq_dev], "gradient_cart2int": [gr_dev], "Hessian_int2cart": [gq_dev, Hq_dev],
                       "Hessian_cart2int": [gr_dev, Hr_dev]}[what]
                hin = [t[:Be].cpu().pin_memory() for t in src]
                oshape = {"gradient_int2cart": (Be, cd), "gradient_cart2int": (Be, n), "Hessian_int2cart": (Be, cd, cd),
                          "Hessian_cart2int": (Be, n, n)}[what]
                houts = [pin(*oshape)]
                ins = [r_host[:Be]] + hin
                fn = getattr(ic, what)
                call = lambda: fn(r_host[:Be], *hin, out=houts[0])
                api = "IntCoordSet.%s with pinned host tensors (copies, kernel, copy back inside the call)" % what
            e2e_steps = max(3, min(steps, 10)) if headline else 3
            for _ in range(2):
                call()
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                call()                     # returns after the results have landed in the host tensors
            torch.cuda.synchronize()
            dt = self.max_over_ranks(time.perf_counter() - t0)
            h2d, d2h = int(sum(t.numel() for t in ins) * 8), int(sum(o.numel() for o in houts) * 8)
            e2e = {"value": self.world * Be * e2e_steps / dt, "unit": "geometries/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "batch_per_gpu": Be, "api": api}
            if headline or self.world > 1:
                pc = self.pcie_ceiling()
                floor_s = h2d / (pc["h2d_GBps_per_gpu"] * 1e9) + d2h / (pc["d2h_GBps_per_gpu"] * 1e9)
                # copies in the two directions overlap (full duplex): the ceiling is the slower direction alone
                duplex_s = max(h2d / (pc["h2d_GBps_per_gpu"] * 1e9), d2h / (pc["d2h_GBps_per_gpu"] * 1e9))
                e2e["copy_ceiling"] = dict(pc, ceiling_geometries_per_s=self.world * Be / duplex_s,
                                           frac_of_ceiling=(self.world * Be * e2e_steps / dt) / (self.world * Be / duplex_s),
                                           serial_copy_geometries_per_s=self.world * Be / floor_s)

        rec = None
        if self.rank == 0:
            peak, peak_src = measured_peak()
            avg_step_s = (sum(step_ms) / steps) * 1e-3
            achieved = bytes_per_geom * B / avg_step_s / 1e9
            kernel = max(kernels, key=kernels.get) if kernels else "none"
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(key) if B == DEFAULT_BATCH[key] else None, "peak_source": peak_src,
                    "algorithmic_bytes_per_geometry": bytes_per_geom, "kernel": kernel, "kernels_launched": kernels,
                    "launch_ms": avg_step_s * 1e3 / launches_per_step}
            if key in FLOPS:
                fl = FLOPS[key](cd, n)
                fpeak = measured_fp64_peak(device)
                tf = fl * B / avg_step_s / 1e12
                roof = {"bound": "tensor", "achieved": tf, "peak": fpeak["burst"], "unit": "TFLOP/s", "frac": tf / fpeak["burst"], "traffic": None,
                        "peak_source": "FP64 DGEMM 8192^3 via cuBLAS (torch.matmul), measured in this run: best of 5 (burst); "
                                       "%.1f TFLOP/s back to back (sustained)" % fpeak["sustained"],
                        "peak_sustained": fpeak["sustained"], "frac_of_sustained": tf / fpeak["sustained"],
                        "dense_equivalent_flops_per_geometry": fl, "kernel": kernel, "kernels_launched": kernels,
                        "launch_ms": avg_step_s * 1e3,
                        "note": "achieved = DENSE-EQUIVALENT flops (the reference algorithm on dense operands, SURVEY 8d, LAPACK counts) / time. "
                                "The kernels use J's structural sparsity (<= 12 nonzeros per row of 90) and execute several times fewer "
                                "flops than that, so this is a rate of useful work, not pipe utilisation; the ncu pipe numbers are in profiles/.",
                        "hbm_GBps": achieved, "hbm_frac": achieved / peak,
                        "algorithmic_bytes_per_geometry": bytes_per_geom}
            rec = {"metric": "geometries/sec (%s)" % what, "value": value, "unit": "geometries/s", "n_gpus": self.world, "steps": steps,
                   "warmup": warmup, "ms_per_step": total_ms_max / steps, "dtype": "f64",
                   "config": {"workload": "%s: %s, %d atoms, %d internal coordinates, per-GPU batch %d%s"
                                          % (key, w.name, w.natoms, n, B, " in chunks of %d" % K_CHUNK if what == "q+J+K" and B > K_CHUNK else ""),
                              "timing": "CUDA events on the launching stream; " + ("L2 flushed between steps" if need_flush else
                                        "inputs+outputs per step exceed the 126 MB L2 (no explicit flush)"),
                              "parallelism": "batch sharded over %d GPU(s), no collective in the data path" % self.world},
                   "roofline": roof, "clocks": sampler.summary(), "gpu_launches": int(launches),
                   "first_call_s": first_call_s,
                   "jit": {"compiled": jit1["compiled"] - jit0["compiled"], "cache_hits": jit1["cache_hits"] - jit0["cache_hits"],
                           "failed": jit1["failed"] - jit0["failed"], "compile_s": jit1["compile_s"] - jit0["compile_s"]}}
            if "cart2int" in what:
                rec["factorisation_failures"] = int(factor_failures)
            if e2e:
                rec["e2e"] = e2e
            if not args.no_cpu_baseline and self.world == 1:
                if self.cpu is None:
                    self.cpu = CpuBaseline()
                rec["cpu_baseline"] = self.cpu.run(wkey, what, per_core_sample(key, headline))
        del r_dev, gq_dev, gr_dev, Hq_dev, Hr_dev
        torch.cuda.empty_cache()
        return rec

    def close(self):
        if self.cpu is not None:
            self.cpu.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="geometries per GPU per step (default: BASELINE size)")
    ap.add_argument("--extras", default=None, help="all | none | comma list of workloads measured into `extra` "
                                                   "(default: all BASELINE configs at 1 GPU, C5,C4H with more)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    wl = load_workloads_module()
    w = wl.get(WL_KEY.get(args.workload, args.workload))
    what = WORKLOADS[args.workload][0]
    B = args.batch or DEFAULT_BATCH[args.workload]
    metric = "geometries/sec (%s)" % what
    if args.impl == "reference":
        config = {"workload": "%s: %s, %d atoms, per-GPU batch %d" % (args.workload, w.name, w.natoms, B),
                  "parallelism": "one single-threaded process per host core"}
        run_reference_arm(args, metric, config)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.extras is None:
        extras = (EXTRAS_1GPU if world == 1 else EXTRAS_NGPU) if args.workload == "C2" and not args.batch else []
    elif args.extras == "all":
        extras = EXTRAS_1GPU
    elif args.extras == "none":
        extras = []
    else:
        extras = [k for k in args.extras.split(",") if k]
    extras = [k for k in extras if k != args.workload]

    bench = Bench(args)
    line = bench.measure(args.workload, B, args.steps, args.warmup, True)
    extra = {}
    for k in extras:
        # extras: fewer steps (each step is already milliseconds to seconds of device time)
        steps = max(3, min(args.steps, 5 if k in ("C3K", "C4Hc", "C4H", "C4gc") else 10))
        rec = bench.measure(k, DEFAULT_BATCH[k], steps, max(3, min(args.warmup, 3)), False)
        if rec is not None:
            extra[k] = rec
    if bench.rank == 0:
        out = {"metric": line["metric"], "value": line["value"], "unit": "geometries/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": line["config"], "impl": "ours",
               "roofline": line["roofline"], "clocks": line["clocks"], "gpu_launches": line["gpu_launches"],
               "first_call_s": line["first_call_s"], "jit": line["jit"]}
        for k in ("e2e", "cpu_baseline", "factorisation_failures"):
            if k in line:
                out[k] = line[k]
        if extra:
            out["extra"] = extra
        print(json.dumps(out))
    bench.close()


if __name__ == "__main__":
    main()
