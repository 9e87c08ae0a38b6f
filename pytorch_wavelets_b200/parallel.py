"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed).

Every (n, c) plane is transformed independently (depthwise filters, no halo between planes), so the path shards
over N with no collective inside the transform.  ``shard_batch`` gives each rank its contiguous slice of the batch;
``gather_pyramid`` is the optional single collective at the end (one all_gather per output tensor, NCCL on GPUs)
for callers that want the reference's full-batch return value on every rank.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous, balanced [start, stop) of a batch of n items for ``rank`` of ``world``."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(x, group=None):
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    a, b = shard_bounds(x.shape[0], world, rank)
    return x[a:b]


def _gather_tensor(t, n_total, group):
    world = dist.get_world_size(group)
    if t.dim() == 0:
        return t
    sizes = [b - a for a, b in (shard_bounds(n_total, world, r) for r in range(world))]
    m = max(sizes)
    t = t.contiguous()
    if t.shape[0] < m:  # uneven shards: pad to the largest so every rank contributes equal-size buffers
        t = torch.cat((t, t.new_zeros((m - t.shape[0],) + tuple(t.shape[1:]))), dim=0)
    parts = [t.new_empty(t.shape) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def gather_pyramid(out, n_total, group=None):
    """All-gather a transform's output structure ((yl, [yh...]) or a tensor) along the batch dimension.
    Band-pass tensors must be batch-major (the default o_dim / ri_dim)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return out
    if isinstance(out, torch.Tensor):
        return _gather_tensor(out, n_total, group)
    if out is None:
        return None
    return type(out)(gather_pyramid(o, n_total, group) for o in out)
