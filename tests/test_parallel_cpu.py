"""CPU, 2 gloo ranks: the N>1 host logic (batch sharding + the optional end-of-transform gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorch_wavelets_b200 import parallel


def test_shard_bounds_cover_and_balance():
    for n in (1, 2, 7, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        x = torch.randn(n, 3, 8, 8)
        xs = parallel.shard_batch(x)
        a, b = parallel.shard_bounds(n, world, rank)
        assert torch.equal(xs, x[a:b])
        # a stand-in "transform" with the output structure of DWTForward / DTCWTForward (incl. a 0-dim placeholder)
        out = (xs.mean(dim=(2, 3), keepdim=True), [xs[:, :, None] * 2.0, xs.new_zeros([]), xs[:, :, None, ::2, ::2, None] * 3.0])
        full = parallel.gather_pyramid(out, n)
        assert torch.equal(full[0], x.mean(dim=(2, 3), keepdim=True))
        assert torch.equal(full[1][0], x[:, :, None] * 2.0)
        assert full[1][1].dim() == 0
        assert torch.equal(full[1][2], x[:, :, None, ::2, ::2, None] * 3.0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [4, 5])
def test_two_rank_shard_and_gather(n):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
