"""Host-side helpers for the multi-GPU modes. torch.distributed is only plumbing here: it carries
64-byte IPC handles / a 128-byte NCCL id once at start-up and provides barriers for benchmarks; the per-iteration
exchange of the sharded ICP happens inside the CUDA kernel over peer-mapped memory."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous index range [lo, hi) of rank's shard: the first n % world ranks get one extra point."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def dist_all_gather_bytes(b: bytes) -> list:
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, b)
    return out


def dist_broadcast_bytes(b):
    import torch.distributed as dist
    box = [b]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def connect(ctx, nccl_baseline=False):
    """Create this rank's communicator on `ctx` using the default torch.distributed process group."""
    import torch.distributed as dist
    ctx.comm_init(dist.get_rank(), dist.get_world_size(), dist_all_gather_bytes, dist_broadcast_bytes, nccl_baseline=nccl_baseline)
