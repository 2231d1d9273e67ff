#!/usr/bin/env python
"""bench.py - geometries/s of the internal-coordinate hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic geometries. At N=1 the
workload is BASELINE.json configs[1] (C2H4, 12 SAS internal coordinates, q + Wilson-B on 10^6
geometries); with N>1 every rank owns its own 10^6-geometry shard (weak scaling, no collective
in the data path). One JSON line is printed by rank 0; see DESIGN.md "Measurement".
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
    "C5x": ("SAPSet value+Jacobian on given coordinates", lambda cd, n, ns: 8 * (n + ns + ns * n)),
    "C5K": ("SAPSet Jacobian2nd on given coordinates", lambda cd, n, ns: 8 * (n + ns * n * n)),
}
DEFAULT_BATCH = {"C1": 10 ** 4, "C2": 10 ** 6, "C3": 10 ** 5, "C3K": 8192, "C4J": 5 * 10 ** 4, "C4g": 10 ** 5, "C4gc": 10 ** 5, "C4H": 10 ** 5,
                 "C4Hc": 10 ** 5, "C5": 12500000, "C5x": 12500000, "C5K": 12500000}
# dense flops per geometry of the contraction-bound transforms
FLOPS = {
    "C4H": lambda cd, n: 2 * n * n * cd + 2 * cd * n * cd + 2 * n * cd * cd,
    "C4Hc": lambda cd, n: 2 * n * n * cd + (n ** 3) // 3 + 2 * (n ** 3) // 3 + 2 * n ** 3 + 2 * n * n * cd
                          + 2 * cd * cd * n + 2 * n * cd * n + 2 * n * cd * cd,
    "C4gc": lambda cd, n: 2 * n * n * cd + (n ** 3) // 3 + 2 * (n ** 3) // 3 + 2 * n ** 3 + 2 * n * n * cd,
}


def measured_fp64_peak(device):
    """cuBLAS DGEMM 4096^3 through torch.matmul, best of 5 (TFLOP/s): the FP64 roofline denominator"""
    import torch
    a = torch.randn(4096, 4096, dtype=torch.float64, device=device)
    b = torch.randn(4096, 4096, dtype=torch.float64, device=device)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


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


# ------------------------------------------------------------------------------------------
# CPU baseline: the reference's own LibTorch implementation (oracle/_ref), one process per core
# ------------------------------------------------------------------------------------------
_WORKER = {}


def _cpu_init(kind, ic_text, what, sap_text, sap_dims):
    """per-process setup: build the reference objects once (table construction is not timed)"""
    if kind == "reference":
        from oracle import ref_lib
        with tempfile.NamedTemporaryFile("w", suffix=".def", delete=False) as f:
            f.write(ic_text)
        s = ref_lib.RefIntCoordSet("default", f.name)
        sap = None
        if "SAP" in what:
            with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f2:
                f2.write(sap_text)
            sap = ref_lib.RefSAPSet(f2.name, 0, sap_dims)
    else:
        from oracle import tchem_oracle as O
        s = O.IntCoordSet("default", ic_text)
        sap = O.SAPSet(sap_text, 0, sap_dims) if "SAP" in what else None
    _WORKER.update(s=s, sap=sap, what=what)


def _cpu_worker(r):
    import time as _t
    s, sap, what = _WORKER["s"], _WORKER["sap"], _WORKER["what"]
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
    numpy port when that library is absent) on one process per host core."""

    def __init__(self, w, what, cores=None):
        import multiprocessing as mp
        from oracle import ref_lib
        self.kind = "reference" if ref_lib.available() else "port"
        try:
            avail = len(os.sched_getaffinity(0))
        except Exception:
            avail = os.cpu_count() or 1
        self.cores = 1 if self.kind == "port" else min(cores or avail, 128)
        self.w, self.what = w, what
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_init,
                                                  initargs=(self.kind, w.ic_text, what, w.sap_text, w.sap_dims))
        self.pool.map(_cpu_worker, [w.geometries(2, seed=1)] * self.cores)     # warm-up / first-call allocations

    def run(self, per_core, seed=4242):
        r = self.w.geometries(per_core * self.cores, seed=seed)
        jobs = [r[i * per_core:(i + 1) * per_core] for i in range(self.cores)]
        t0 = time.perf_counter()
        times = self.pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        # SAPSet on given coordinates: the workers time the SAPSet call alone (the coordinates are the caller's)
        span = max(times) if self.what.startswith("SAPSet") else wall
        return {"value": per_core * self.cores / span, "unit": "geometries/s", "cores": self.cores, "kind": self.kind,
                "sample": "%d geometries (%d per process, %d single-threaded processes), %s; wall %.2f s, slowest process %.2f s"
                          % (per_core * self.cores, per_core, self.cores, self.what, wall, max(times))}

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(w, what, per_core, cores=None):
    cb = CpuBaseline(w, what, cores)
    try:
        return cb.run(per_core)
    finally:
        cb.close()


def per_core_sample(key):
    # sized for roughly 10-20 s of CPU work per process with the reference's measured per-geometry cost
    return {"C1": 20000, "C2": 4000, "C3": 1500, "C3K": 24, "C4J": 1000, "C4g": 1000, "C4gc": 1000, "C4H": 16, "C4Hc": 16, "C5": 20000, "C5x": 20000, "C5K": 20000}[key]


# ------------------------------------------------------------------------------------------
def run_reference_arm(args, w, what, metric, config):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_core = max(4, per_core_sample(args.workload) // 8)      # ~1-2 s of CPU work per step
    vals = []
    base = None
    cb = CpuBaseline(w, what)
    for i in range(args.warmup + args.steps):
        base = cb.run(per_core, seed=100 + i)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="geometries per GPU per step (default: BASELINE size)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    import importlib.util
    spec = importlib.util.spec_from_file_location("_wl", os.path.join(ROOT, "torch_chemistry_b200", "workloads.py"))
    wl = importlib.util.module_from_spec(spec)
    sys.modules["_wl"] = wl
    spec.loader.exec_module(wl)
    wkey = {"C3K": "C3", "C5x": "C5", "C5K": "C5", "C4J": "C4", "C4g": "C4", "C4gc": "C4", "C4H": "C4", "C4Hc": "C4"}.get(args.workload, args.workload)
    w = wl.get(wkey)
    what, bytes_fn = WORKLOADS[args.workload]
    B = args.batch or DEFAULT_BATCH[args.workload]
    metric = "geometries/sec (%s)" % what
    config = {"workload": "%s: %s, %d atoms, per-GPU batch %d" % (args.workload, w.name, w.natoms, B),
              "timing": "CUDA events on the launching stream; inputs+outputs per step exceed the 126 MB L2 "
                        "(no explicit flush)" if B * bytes_fn(w.cartdim, 1, 1) > 126e6 else
                        "CUDA events on the launching stream; L2 flushed between steps",
              "parallelism": "batch sharded over %d GPU(s), no collective in the data path" % args.gpus}

    if args.impl == "reference":
        run_reference_arm(args, w, what, metric, config)
        return

    import torch
    import torch.distributed as dist
    import torch_chemistry_b200 as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    ic = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms, device=local)
    sap = T.SAPSet.from_text(w.sap_text, 0, w.sap_dims, device=local) if w.sap_text else None
    n, cd = ic.intdim, ic.cartdim
    ns = sap.nterms if sap else 0
    bytes_per_geom = bytes_fn(cd, n, ns)

    # synthetic shard of this rank: same generator, rank-dependent seed (shard-invariant definition)
    r_host = torch.from_numpy(w.geometries(B, seed=w.seed + 1000 * rank)).pin_memory()
    r_dev = r_host.to(device)
    f64 = dict(dtype=torch.float64, device=device)
    gq_dev = torch.randn(B, n, **f64) if what in ("gradient_int2cart", "Hessian_int2cart") else None
    gr_dev = torch.randn(B, cd, **f64) if what in ("gradient_cart2int", "Hessian_cart2int") else None
    Hq_dev = Hr_dev = None
    if what == "Hessian_int2cart":
        Hq_dev = torch.randn(B, n, n, **f64)
        Hq_dev = 0.5 * (Hq_dev + Hq_dev.transpose(1, 2)).contiguous()
    if what == "Hessian_cart2int":
        Hr_dev = torch.randn(B, cd, cd, **f64)
        Hr_dev = 0.5 * (Hr_dev + Hr_dev.transpose(1, 2)).contiguous()

    if what == "q+J":
        outs = [torch.empty(B, n, **f64), torch.empty(B, n, cd, **f64)]
        step = lambda: ic.compute_IC_J(r_dev, out=outs)
    elif what == "q+J+K":
        outs = [torch.empty(B, n, **f64), torch.empty(B, n, cd, **f64), torch.empty(B, n, cd, cd, **f64)]
        step = lambda: ic.compute_IC_J_K(r_dev, out=outs)
    elif what == "gradient_int2cart":
        step = lambda: ic.gradient_int2cart(r_dev, gq_dev)
    elif what == "gradient_cart2int":
        step = lambda: ic.gradient_cart2int(r_dev, gr_dev)
    elif what == "Hessian_int2cart":
        step = lambda: ic.Hessian_int2cart(r_dev, gq_dev, Hq_dev)
    elif what == "Hessian_cart2int":
        step = lambda: ic.Hessian_cart2int(r_dev, gr_dev, Hr_dev)
    elif args.workload == "C5K":
        import ctypes as C
        x_dev = ic(r_dev).contiguous()
        K2 = torch.empty(B, ns, n, n, **f64)
        step = lambda: T._capi.check(T._lib.tchem_sap_jacobian2nd(sap._handle, C.c_void_p(x_dev.data_ptr()), B, C.c_void_p(K2.data_ptr()),
                                                                  C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    elif args.workload == "C5x":
        x_dev = ic(r_dev).contiguous()                      # the coordinates a caller would bring
        souts = sap.eval_cat(x_dev)
        step = lambda: sap.eval_cat(x_dev, out=souts)
    else:
        step = lambda: sap.chained(ic, r_dev)
    need_flush = B * bytes_per_geom <= 126e6
    flush_buf = torch.empty(160 * 1024 * 1024 // 8, **f64) if need_flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = T.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        if need_flush:
            flush_buf.zero_()
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    launches = T.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * B * args.steps / (total_ms_max * 1e-3)

    # ---- e2e: host buffers through the public API, copies inside the timed region --------------
    e2e = None
    if not args.no_e2e and args.workload not in ("C5x", "C5K") and (what in ("q+J", "q+J+K") or sap is not None):
        pin = lambda *s: torch.empty(s, dtype=torch.float64).pin_memory()
        if sap is not None:
            houts = [pin(B, ns), pin(B, ns, n)]
            call = lambda: sap.chained(ic, r_host, out=houts)
        else:
            houts = [pin(B, n), pin(B, n, cd)] + ([pin(B, n, cd, cd)] if what == "q+J+K" else [])
            call = (lambda: ic.compute_IC_J(r_host, out=houts)) if what == "q+J" else (lambda: ic.compute_IC_J_K(r_host, out=houts))
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()                     # returns after the results have landed in the host tensors
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * e2e_steps / float(tt.item()), "unit": "geometries/s",
               "h2d_bytes_per_step": int(B * cd * 8), "d2h_bytes_per_step": int(sum(o.numel() for o in houts) * 8),
               "steps": e2e_steps, "api": ("SAPSet.chained(IntCoordSet, r) with pinned host tensors -> tchem_ic_sap_eval_host" if sap is not None
                       else "IntCoordSet.compute_IC_J(r) with pinned host tensors -> tchem_ic_eval_host")}

    if rank == 0:
        peak, peak_src = measured_peak()
        avg_launch_s = (sum(step_ms) / args.steps) * 1e-3
        achieved = bytes_per_geom * B / avg_launch_s / 1e9
        bulk = os.environ.get("TCHEM_JIT_BULK", "1") != "0"
        chain_name = "tchem_jit_chain_bulk" if bulk else "tchem_jit_chain"
        kernel = ("tchem_jit_sap_hess (SAPSet-specialised, NVRTC, one TMA bulk store per tile)" if args.workload == "C5K" and B >= 16384
                  else "sap_hess_kernel" if args.workload == "C5K"
                  else chain_name + " (SAPSet-specialised, NVRTC)" if args.workload == "C5x" and T._lib.tchem_ic_sap_uses_specialised(sap._handle, None, B)
                  else chain_name + " (pair-specialised, NVRTC)" if sap is not None and T._lib.tchem_ic_sap_uses_specialised(sap._handle, ic._handle, B)
                  else "sap_eval_kernel" if sap is not None else "grad_int2cart_kernel" if what == "gradient_int2cart"
                  else "dense_kernel" if args.workload in FLOPS
                  else ("tchem_jit_qJ_bulk (table-specialised, NVRTC, TMA tensor stores)" if os.environ.get("TCHEM_JIT_BULK", "1") != "0" and args.workload == "C2"
                        else "tchem_jit_qJ (table-specialised, NVRTC)") if what == "q+J" and T._lib.tchem_ic_uses_specialised(ic._handle, B)
                  else "ic_eval_kernel")
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload) if B == DEFAULT_BATCH[args.workload] else None, "peak_source": peak_src,
                "algorithmic_bytes_per_geometry": bytes_per_geom, "kernel": kernel, "launch_ms": avg_launch_s * 1e3}
        if args.workload in FLOPS:
            fl = FLOPS[args.workload](cd, n)
            fpeak = measured_fp64_peak(device)
            tf = fl * B / avg_launch_s / 1e12
            roof = {"bound": "tensor", "achieved": tf, "peak": fpeak, "unit": "TFLOP/s", "frac": tf / fpeak, "traffic": None,
                    "peak_source": "FP64 DGEMM 4096^3 via cuBLAS (torch.matmul), measured in this run",
                    "dense_flops_per_geometry": fl, "kernel": kernel, "launch_ms": avg_launch_s * 1e3,
                    "note": "achieved = dense-equivalent (algorithmic, SURVEY 8d) flops / time; the kernel skips DMMA k-steps in which J is structurally zero, so it issues fewer flops than that",
                    "hbm_GBps": achieved, "hbm_frac": achieved / peak}
        line = {"metric": metric, "value": value, "unit": "geometries/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "impl": "ours", "roofline": roof,
                "clocks": sampler.summary(), "gpu_launches": int(launches)}
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(w, what, per_core_sample(args.workload))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
