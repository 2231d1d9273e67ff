"""Multi-GPU plumbing on the CPU: world_size-2 (and 3) gloo process groups exercise the batch
partition and the gather of ragged output shards (the device kernels themselves are covered by
the -m gpu tests; sharding adds no arithmetic)."""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

_spec = importlib.util.spec_from_file_location("_tchem_sharding", os.path.join(ROOT, "torch_chemistry_b200", "sharding.py"))
sharding = importlib.util.module_from_spec(_spec)
sys.modules["_tchem_sharding"] = sharding
_spec.loader.exec_module(sharding)


def test_shard_bounds_cover_the_batch_exactly():
    for batch in (0, 1, 7, 32, 10 ** 6, 10 ** 8 + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_bounds(batch, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == batch
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(batch * 6, dtype=torch.float64).reshape(batch, 2, 3)

        def local(begin, end):                 # stands in for the kernel: any per-geometry function
            return full[begin:end] * 2.0 + 1.0

        shard = sharding.evaluate_sharded(local, batch)
        out = sharding.gather_shards(shard, batch, dst=0)
        if rank == 0:
            np.save(os.path.join(tmp, "gathered.npy"), out.numpy())
        else:
            assert out is None
        # a wrong shard size is refused on every rank
        try:
            sharding.gather_shards(torch.zeros(shard.shape[0] + 1, 2, 3, dtype=torch.float64), batch)
            raise AssertionError("size check missing")
        except ValueError:
            pass
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 11), (2, 64), (3, 10)])
def test_gather_of_ragged_shards_gloo(tmp_path, world, batch):
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    want = np.arange(batch * 6, dtype=np.float64).reshape(batch, 2, 3) * 2.0 + 1.0
    np.testing.assert_array_equal(got, want)
