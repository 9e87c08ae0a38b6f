"""-m gpu, needs >= 2 GPUs (skipped otherwise): the sharded path on real devices -- each rank transforms its shard, the
pyramid is all-gathered over NCCL (through the C ABI's communicator and through torch.distributed) and must equal the
single-GPU transform of the whole batch bit for bit."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n):
    import torch.distributed as dist
    import pytorch_wavelets_b200 as pw
    from pytorch_wavelets_b200 import parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    comm = None
    try:
        torch.manual_seed(0)
        x = torch.randn(n, 3, 96, 128)
        xs = parallel.shard_batch(x).to(dev)
        comm = parallel.Communicator.from_torch_distributed()
        for make in (lambda: pw.DWTForward(J=3, wave='db4', mode='symmetric'), lambda: pw.DTCWTForward(J=3)):
            f = make().to(dev)
            with torch.no_grad():
                out = f(xs)
                full_c = parallel.gather_pyramid(out, n, comm=comm)          # C-ABI NCCL communicator
                full_t = parallel.gather_pyramid(out, n)                     # torch.distributed (NCCL)
                ref = f(x.to(dev))                                           # the whole batch on one GPU
            for full in (full_c, full_t):
                assert torch.equal(full[0], ref[0])
                assert len(full[1]) == len(ref[1])
                for a, b in zip(full[1], ref[1]):
                    assert torch.equal(a, b)
        torch.cuda.synchronize()
    finally:
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [4, 5])
def test_two_rank_nccl_gather_equals_single_gpu(n):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
