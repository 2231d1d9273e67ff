"""Batch sharding over the GPUs of one node.

Every geometry is independent, so the batch axis is cut into contiguous slices
[rank * B / N, (rank + 1) * B / N) - one process per GPU, the IC / SAP tables replicated, NO
collective in the evaluate loop. Only when the caller wants the results on one device are the
output shards gathered (NCCL over NVLink through torch.distributed; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world_size, rank):
    """[begin, end) of `rank`'s contiguous slice; sizes differ by at most one geometry."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(batch), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_sizes(batch, world_size):
    return [shard_bounds(batch, world_size, r)[1] - shard_bounds(batch, world_size, r)[0] for r in range(world_size)]


def evaluate_sharded(fn, batch, world_size=None, rank=None):
    """Runs fn(begin, end) on this rank's slice of a batch of `batch` geometries."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    begin, end = shard_bounds(batch, world_size, rank)
    return fn(begin, end)


def gather_shards(shard, batch, dst=0, group=None):
    """Collects the per-rank output shards (leading axis = this rank's slice of the batch) on rank
    `dst` in batch order. Returns the [batch, ...] tensor on `dst`, None elsewhere. Shards may be
    ragged (batch not divisible by the world size): point-to-point sends into the right rows of
    the destination tensor, no padding and no staging copy."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(batch, world)
    if shard.shape[0] != sizes[rank]:
        raise ValueError("shard holds %d geometries, this rank owns %d" % (shard.shape[0], sizes[rank]))
    shard = shard.contiguous()
    if rank == dst:
        out = torch.empty((batch,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        ops = []
        for src in range(world):
            b, e = shard_bounds(batch, world, src)
            if src == dst:
                out[b:e].copy_(shard)
            elif e > b:
                ops.append(dist.P2POp(dist.irecv, out[b:e], src, group))
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        return out
    if shard.shape[0] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, shard, dst, group)]):
            req.wait()
    return None
