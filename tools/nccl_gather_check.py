#!/usr/bin/env python
"""Sharded evaluation over the GPUs of one node + NCCL gather of the result shards on rank 0 (SURVEY.md section 5 / 8e:
"kernel vs gather"), checked bit for bit against a single-GPU run.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/nccl_gather_check.py

Workloads: C2 q+J (10^6 geometries in total, 1.82 GB of J), C3 q+J+K (8 192 geometries in total, 12.7 GB of K - the
case where the gather, not the kernel, bounds the job). One JSON line per workload on rank 0: device time of the sharded
kernels (max over ranks), device time of the gather, GB/s into the root against the ~900 GB/s NVLink-5 ingress of one GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_chemistry_b200 as T  # noqa: E402
from torch_chemistry_b200 import sharding  # noqa: E402


def timed(fn, reps=3):
    """best device time (ms) of fn over `reps` runs, all ranks in step; returns (ms, last result)"""
    best, out = None, None
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
    return best, out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dev = torch.device("cuda", torch.cuda.current_device())
    for key, B, want_K in (("C2", 1000003, False), ("C3", 8192, True)):
        w = T.workloads.get(key)
        s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms, device=dev)
        r = torch.from_numpy(w.geometries(B, seed=17))     # same bits on every rank
        b, e = sharding.shard_bounds(B, world, rank)
        rd = r[b:e].to(dev)
        run = (lambda: s.compute_IC_J_K(rd)) if want_K else (lambda: s.compute_IC_J(rd))
        run()
        kernel_ms, outs = timed(run)
        big = outs[-1]                                      # J or K: the array that matters
        sharding.gather_shards(outs[0], B, dst=0)           # warm-up: communicator set-up
        gather_ms, full = timed(lambda: sharding.gather_shards(big, B, dst=0))
        if rank == 0:
            bytes_in = full.numel() * 8 * (B - (e - b)) / B
            ok = None
            if not want_K or world <= 8:
                # reference: the same batch on one GPU, in slices that fit beside the gathered tensor
                ok = True
                # (slices of >= 16 384 geometries each: q+J then runs the same specialised kernel as the sharded call - a
                # 3-geometry remainder would be served by the interpreter kernel, whose last bits may differ)
                step = 1024 if want_K else (B + 3) // 4
                for c0 in range(0, B, step):
                    ref = (s.compute_IC_J_K if want_K else s.compute_IC_J)(r[c0:c0 + step].to(dev))[-1]
                    ok = ok and torch.equal(full[c0:c0 + step], ref)
            print(json.dumps({"workload": "%s %s" % (key, "q+J+K" if want_K else "q+J"), "n_gpus": world, "geometries": B,
                              "kernel_ms_sharded": kernel_ms, "gather_ms": gather_ms, "gathered_GB": full.numel() * 8 / 1e9,
                              "root_ingress_GBps": bytes_in / 1e9 / (gather_ms * 1e-3), "root_ingress_bound_GBps": 900.0,
                              "gather_over_kernel": gather_ms / kernel_ms,
                              "bit_identical_to_one_gpu": ok}), flush=True)
            if ok is False:
                sys.exit(1)
        del outs, big, full
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
