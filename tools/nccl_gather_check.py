#!/usr/bin/env python
"""Sharded q+J over the GPUs of one node + NCCL gather on rank 0, checked against a single-GPU run.
torchrun --nproc-per-node N tools/nccl_gather_check.py   (N = 2, 4, 8)"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_chemistry_b200 as T  # noqa: E402
from torch_chemistry_b200 import sharding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    w = T.workloads.get("C2")
    s = T.IntCoordSet.from_text("default", w.ic_text, n_atoms=w.natoms, device=torch.device("cuda", torch.cuda.current_device()))
    B = 200003                                         # ragged over the ranks
    r = torch.from_numpy(w.geometries(B, seed=17))     # same bits on every rank
    b, e = sharding.shard_bounds(B, world, rank)
    q, J = s.compute_IC_J(r[b:e].cuda())
    sharding.gather_shards(q, B, dst=0)                 # warm-up: communicator set-up
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    Jall = sharding.gather_shards(J, B, dst=0)
    qall = sharding.gather_shards(q, B, dst=0)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        qref, Jref = s.compute_IC_J(r.cuda())
        ok = torch.equal(Jall, Jref) and torch.equal(qall, qref)
        gb = (Jall.numel() + qall.numel()) * 8 * (world - 1) / world / 1e9
        print("nccl gather over %d GPUs: %s, %.2f GB received in %.1f ms (%.0f GB/s)" % (world, "bit-identical to the 1-GPU result" if ok else "MISMATCH", gb, dt * 1e3, gb / dt))
        if not ok:
            sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
