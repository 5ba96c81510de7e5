"""Multi-GPU plumbing for the batch-parallel primitives.

Every NTT / dyadic multiply / keyswitch item is independent, so N GPUs take N
contiguous slices of the batch and never exchange data on the data path (the
reference scales the same way: NUM_DEV threads popping one queue,
host/src/fpga.cpp:1646-1673).  The only collective is the one-off replication
of the read-only tables (twiddles, switch keys) from rank 0.
"""


def shard(batch, world, rank):
    """Contiguous slice [start, start+count) of `batch` items owned by `rank`:
    the first batch % world ranks get one extra item."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def replicate(tensors, src=0, group=None):
    """Broadcast read-only tables from `src` to every rank (NCCL over NVLink on
    GPUs, gloo in the CPU tests).  `tensors`: list of same-shaped-on-all-ranks
    tensors, filled on `src`."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.broadcast(t, src, group=group)
    return tensors
