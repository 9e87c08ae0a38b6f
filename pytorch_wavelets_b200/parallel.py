"""Batch sharding across the GPUs of one box (one process per GPU).

Every (n, c) plane is transformed independently (depthwise filters, no halo between planes), so the path shards
over N with no collective inside the transform.  ``shard_batch`` gives each rank its contiguous slice of the batch;
``gather_pyramid`` is the single collective at the end (north_star: "a single NCCL gather at the end"): one
all-gather per output tensor along dim 0, enqueued on the compute stream right behind the last level.

Two transports:
  * ``Communicator`` -- the C ABI's own NCCL communicator (``b200w_comm_init`` / ``b200w_allgather`` in
    include/b200wave.h); ``torch.distributed`` is used only to hand the 128-byte NCCL id to the other ranks.
  * ``torch.distributed.all_gather`` on the default / given process group (NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes

import torch
import torch.distributed as dist

from pytorch_wavelets_b200 import _ffi


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


class Communicator(object):
    """One NCCL communicator for this process's current CUDA device, owned through the C ABI.

    ``Communicator.from_torch_distributed()`` creates it collectively on every rank of an initialised
    ``torch.distributed`` job (any backend: the process group only carries the 128-byte NCCL id)."""

    def __init__(self, rank, world, unique_id):
        self.rank, self.world = int(rank), int(world)
        self._h = ctypes.c_void_p()
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(unique_id))
        rc = _ffi.lib().b200w_comm_init(ctypes.byref(self._h), self.rank, self.world, buf)
        self._check(rc, 'b200w_comm_init')

    @staticmethod
    def _check(rc, what):
        if rc == 0:
            return
        if rc == -6:
            raise NotImplementedError('%s: NCCL is not available in this process' % what)
        raise _ffi.B200WaveError('%s failed (%d): %s' % (what, rc, _ffi.lib().b200w_comm_last_error().decode()))

    @staticmethod
    def unique_id():
        buf = (ctypes.c_char * 128)()
        Communicator._check(_ffi.lib().b200w_comm_unique_id(buf), 'b200w_comm_unique_id')
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, group=None):
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(rank, world, box[0])

    def all_gather(self, t):
        """(world * n, ...) tensor whose slice r is rank r's contiguous fp32 CUDA tensor ``t`` (equal shapes)."""
        _ffi.require_cuda_f32(t, 'tensor')
        t = t.contiguous()
        out = t.new_empty((self.world * t.shape[0],) + tuple(t.shape[1:]))
        with torch.cuda.device(t.device):
            rc = _ffi.lib().b200w_allgather(self._h, t.data_ptr(), out.data_ptr(), t.numel(), _ffi.stream_of(t))
        self._check(rc, 'b200w_allgather')
        return out

    def close(self):
        if self._h:
            _ffi.lib().b200w_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _sizes(n_total, world):
    return [b - a for a, b in (shard_bounds(n_total, world, r) for r in range(world))]


def _gather_tensor(t, n_total, group, comm):
    if t.dim() == 0:
        return t
    world = comm.world if comm is not None else dist.get_world_size(group)
    sizes = _sizes(n_total, world)
    m = max(sizes)
    t = t.contiguous()
    if t.shape[0] < m:  # uneven shards: pad to the largest so every rank contributes equal-size buffers
        t = torch.cat((t, t.new_zeros((m - t.shape[0],) + tuple(t.shape[1:]))), dim=0)
    if comm is not None:
        g = comm.all_gather(t)
        if min(sizes) == m:
            return g
        return torch.cat([g[r * m: r * m + n] for r, n in enumerate(sizes)], dim=0)
    parts = [t.new_empty(t.shape) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def gather_pyramid(out, n_total, group=None, comm=None):
    """All-gather a transform's output structure ((yl, [yh...]) or a tensor) along the batch dimension.
    Band-pass tensors must be batch-major (the default o_dim / ri_dim).  ``comm``: a :class:`Communicator`
    (the C ABI's NCCL path); otherwise ``torch.distributed`` on ``group``."""
    if comm is None and (not dist.is_initialized() or dist.get_world_size(group) == 1):
        return out
    if comm is not None and comm.world == 1:
        return out
    if isinstance(out, torch.Tensor):
        return _gather_tensor(out, n_total, group, comm)
    if out is None:
        return None
    return type(out)(gather_pyramid(o, n_total, group, comm) for o in out)
