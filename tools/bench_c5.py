"""BASELINE.json configs[4]: DWTForward J=4 db8 (mode zero) on N=1024, C=16, 2048x2048 fp32, sharded over N across the
GPUs of one box.  The input alone is 275 GB, so every rank streams its shard through the transform in chunks of
--chunk images (synthetic data generated once per rank and reused; outputs are produced and dropped chunk by chunk,
i.e. rank-resident / consumed in place -- SURVEY 8(e)).  One JSON line from rank 0: total Mpix/s over all ranks.

  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/bench_c5.py --n-total 1024
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import parallel

ap = argparse.ArgumentParser()
ap.add_argument('--n-total', type=int, default=1024)
ap.add_argument('--chunk', type=int, default=16)
ap.add_argument('--max-chunks', type=int, default=0, help='time only this many chunks per rank (0 = the whole shard)')
a = ap.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local); dev = torch.device('cuda', local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
lo, hi = parallel.shard_bounds(a.n_total, world, rank)
n_chunks = (hi - lo + a.chunk - 1) // a.chunk
if a.max_chunks: n_chunks = min(n_chunks, a.max_chunks)
torch.manual_seed(100 + rank)
x = torch.randn(a.chunk, 16, 2048, 2048, device=dev)
f = pw.DWTForward(J=4, wave='db8', mode='zero').to(dev)
with torch.no_grad():
    for _ in range(3): f(x)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_chunks): yl, yh = f(x)
    e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
pix = world * n_chunks * x.numel()
if rank == 0:
    print(json.dumps({'config': 'C5 DWT J=4 db8 zero, N=%d sharded over %d GPU(s), chunk %d' % (a.n_total, world, a.chunk),
                      'chunks_per_rank': n_chunks, 'seconds': float(t.item()), 'mpix_s': pix / float(t.item()) / 1e6,
                      'alg_gbps_per_gpu': 8.108 * pix / world / float(t.item()) / 1e9}))
if world > 1: dist.destroy_process_group()
