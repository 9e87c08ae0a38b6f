#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 wavelet filterbank engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N>1: launched by the driver as  python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...)

One "step" = one pass of the hot path over one batch of synthetic input:
    DWTForward(J=3,'db4','symmetric') on randn(128,32,512,512)          (BASELINE.json configs[1])
  + DTCWTForward(J=3,'near_sym_a','qshift_a') on randn(64,3,1024,1024)  (BASELINE.json configs[2])
i.e. the two transforms BASELINE.json's metric names ("Mpixels/sec DWT J=3 db4 + DTCWT J=3 fwd").
`value` = input pixels of both transforms on all ranks / device time (max over ranks), inputs resident
in HBM; per-transform figures are in `parts`.  Inputs (4.3 GB + 0.8 GB) are far larger than the 126 MB L2,
so no explicit flush is needed between iterations (stated in config.l2).
Multi-GPU: every rank transforms its own batch shard of the same size (weak scaling, no data-path
collective; outputs stay rank-resident -- see DESIGN.md section "multi-GPU").

Extra objects on the JSON line (see the task contract): roofline (dominant kernel, algorithmic bytes over
live CUDA-event time), cpu_baseline (oracle port timed on the host cores, bounded sample), e2e (host
buffers through the public nn.Module API, H2D + D2H inside the timed region), clocks, gpu_launches.

--impl reference: times the CPU implementation of the same path (the oracle port, all host threads) on a
bounded sample of the same workload; rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DWT_SHAPE = (128, 32, 512, 512)
DTCWT_SHAPE = (64, 3, 1024, 1024)
METRIC = 'Mpixels/sec DWT J=3 db4 + DTCWT J=3 fwd'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--dwt-batch', type=int, default=DWT_SHAPE[0])
    ap.add_argument('--dtcwt-batch', type=int, default=DTCWT_SHAPE[0])
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample (also the cpu_baseline of the GPU arm)

def cpu_sample(dwt_n=4, dtcwt_n=4, reps=3):
    import numpy as np
    from oracle import oracle as orc
    from pytorch_wavelets_b200 import wavelets
    from pytorch_wavelets_b200.dtcwt import coeffs
    orc.lib()
    w = wavelets.Wavelet('db4')
    h0, h1 = np.array(w.dec_lo[::-1]), np.array(w.dec_hi[::-1])
    h0o, _, h1o, _ = coeffs.biort('near_sym_a')
    q = coeffs.qshift('qshift_a')
    l1 = (h0o[::-1].ravel(), h1o[::-1].ravel())
    qs = tuple(q[i][::-1].ravel() for i in (0, 1, 4, 5))
    rng = np.random.default_rng(0)
    xd = rng.standard_normal((dwt_n,) + DWT_SHAPE[1:]).astype(np.float32)
    xt = rng.standard_normal((dtcwt_n,) + DTCWT_SHAPE[1:]).astype(np.float32)
    orc.dwt_forward(xd[:1], (h0, h1, h0, h1), 3, 'symmetric')  # warm-up (page-in, threads)
    best_d = best_t = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        orc.dwt_forward(xd, (h0, h1, h0, h1), 3, 'symmetric')
        best_d = min(best_d, time.perf_counter() - t)
        t = time.perf_counter()
        orc.dtcwt_forward(xt, l1, qs, 3)
        best_t = min(best_t, time.perf_counter() - t)
    # time the same MIX of work as one GPU step: pixels weighted like the full workload
    pd, pt = float(np.prod(DWT_SHAPE)), float(np.prod(DTCWT_SHAPE))
    rate_d, rate_t = xd.size / best_d, xt.size / best_t
    step_time = pd / rate_d + pt / rate_t
    return {
        'value': (pd + pt) / step_time / 1e6,
        'dwt_mpix_s': rate_d / 1e6,
        'dtcwt_mpix_s': rate_t / 1e6,
        'sample': 'DWT %dx%dx%dx%d + DTCWT %dx%dx%dx%d, best of %d, extrapolated to the full step mix' % (
            (dwt_n,) + DWT_SHAPE[1:] + (dtcwt_n,) + DTCWT_SHAPE[1:] + (reps,)),
        'step_s': step_time,
    }


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault('OMP_NUM_THREADS', str(cores))
    vals = []
    t0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        r = cpu_sample(dwt_n=2, dtcwt_n=2, reps=1)
        if i >= args.warmup:
            vals.append(r)
        if time.perf_counter() - t0 > 150:  # keep the whole arm within a few minutes
            break
    if not vals:
        vals = [r]
    v = sum(x['value'] for x in vals) / len(vals)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v * args.gpus, 'unit': 'Mpix/s', 'n_gpus': args.gpus,
        'steps': len(vals), 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(x['step_s'] for x in vals) / len(vals),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': v, 'unit': 'Mpix/s', 'cores': cores, 'kind': 'port',
                         'sample': vals[-1]['sample'] + '; step time extrapolated from the sample',
                         'dwt_mpix_s': vals[-1]['dwt_mpix_s'], 'dtcwt_mpix_s': vals[-1]['dtcwt_mpix_s']},
        'e2e': {'value': v * args.gpus, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'reference arm = CPU oracle port of the reference algorithm (oracle/wave_oracle.c, OpenMP over '
                'planes, all host threads); the reference is pure Python and cannot travel to the GPU box. '
                'n_gpus>1: value = per-host figure x n (the CPU path does not use GPUs).',
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        'workload': 'DWTForward(J=3,db4,symmetric) %dx32x512x512 + DTCWTForward(J=3,near_sym_a,qshift_a) '
                    '%dx3x1024x1024 fp32 per GPU (BASELINE configs[1]+configs[2])' % (args.dwt_batch, args.dtcwt_batch),
        'per_gpu_pixels': args.dwt_batch * 32 * 512 * 512 + args.dtcwt_batch * 3 * 1024 * 1024,
        'parallelism': 'batch-sharded replicas x%d, outputs rank-resident, no data-path collective' % world,
        'l2': 'inputs (>=5 GB per step) exceed the 126 MB L2; no explicit flush',
    }


# ------------------------------------------------------------------------------------------------------
# clocks sampler (pynvml; falls back to nvidia-smi)

class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.005)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {'sm_mhz': (s[len(s) // 2] if s else None), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


# ------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import pytorch_wavelets_b200 as pw
    from pytorch_wavelets_b200 import _ffi

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the product path)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    _ffi.lib()

    dshape = (args.dwt_batch,) + DWT_SHAPE[1:]
    tshape = (args.dtcwt_batch,) + DTCWT_SHAPE[1:]
    torch.manual_seed(1234 + rank)
    xd = torch.randn(dshape, device=dev)
    xt = torch.randn(tshape, device=dev)
    dwt = pw.DWTForward(J=3, wave='db4', mode='symmetric').to(dev)
    dtc = pw.DTCWTForward(J=3, biort='near_sym_a', qshift='qshift_a').to(dev)
    pix_d, pix_t = xd.numel(), xt.numel()

    def step():
        with torch.no_grad():
            a = dwt(xd)
            b = dtc(xt)
        return a, b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    del out
    barrier()

    # -- timed region: whole step, device events; per-call events through the FFI hook for the roofline
    rec = _ffi.CallRecorder()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    barrier()
    with rec:
        ev[0].record()
        for _ in range(args.steps):
            with torch.no_grad():
                a = dwt(xd)
        ev[1].record()
        for _ in range(args.steps):
            with torch.no_grad():
                b = dtc(xt)
        ev[2].record()
    barrier()
    clocks = sampler.stop()
    t_d = ev[0].elapsed_time(ev[1]) / 1e3
    t_t = ev[1].elapsed_time(ev[2]) / 1e3
    t_total = t_d + t_t
    calls = rec.summary()
    launches = rec.count
    del a, b

    tt = torch.tensor([t_total, t_d, t_t], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_d, t_t = [float(v) for v in tt.tolist()]

    value = world * (pix_d + pix_t) * args.steps / t_total / 1e6
    parts = {
        'dwt_fwd_mpix_s': world * pix_d * args.steps / t_d / 1e6,
        'dtcwt_fwd_mpix_s': world * pix_t * args.steps / t_t / 1e6,
        'dwt_ms': 1e3 * t_d / args.steps, 'dtcwt_ms': 1e3 * t_t / args.steps,
    }

    # -- roofline of the dominant kernel (largest share of the step)
    peak, peak_src = load_peaks()
    alg = {
        'dwt': 4.0 * dshape[0] * dshape[1] * (512 * 512 + 3 * (259 ** 2 + 133 ** 2 + 70 ** 2) + 70 ** 2),
        'dtcwt': 4.0 * tshape[0] * tshape[1] * (1024 * 1024 * (1 + 3 + 0.75 + 0.1875) + 256 * 256),
    }
    roof = None
    if calls:
        top = max(calls.values(), key=lambda c: c['total_ms'])
        ach = top['alg_bytes'] / (top['avg_ms'] / 1e3) / 1e9
        # DRAM bytes of one launch of that kernel, from the committed ncu --set full capture (not measurable live)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
            if top['tag'] in tj and args.dwt_batch == DWT_SHAPE[0]:
                traffic = tj[top['tag']]['dram_bytes_per_launch']
        except Exception:
            pass
        roof = {'bound': 'hbm', 'kernel': top['tag'], 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                'frac': ach / peak, 'traffic': traffic, 'peak_source': peak_src,
                'alg_bytes_per_launch': top['alg_bytes'], 'avg_launch_ms': top['avg_ms'],
                'share_of_step': top['total_ms'] / (1e3 * (t_d + t_t)),
                'whole_transform': {
                    'dwt_fwd_GBps': alg['dwt'] * args.steps / t_d / 1e9, 'dwt_frac': alg['dwt'] * args.steps / t_d / 1e9 / peak,
                    'dtcwt_fwd_GBps': alg['dtcwt'] * args.steps / t_t / 1e9, 'dtcwt_frac': alg['dtcwt'] * args.steps / t_t / 1e9 / peak},
                'levels': {k: {'avg_ms': round(v['avg_ms'], 4), 'GBps': round(v['alg_bytes'] / v['avg_ms'] / 1e6, 1)}
                           for k, v in sorted(calls.items())}}

    # -- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(torch, dist, pw, dev, world, dshape, tshape, min(args.steps, 5))
        except Exception as exc:  # e.g. not enough pinnable host memory for N ranks: report, do not lose the line
            e2e = {'value': None, 'unit': 'Mpix/s', 'error': repr(exc)[:200]}
            if world > 1:
                try:
                    dist.barrier()
                except Exception:
                    pass

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        c = cpu_sample()
        cpu = {'value': c['value'], 'unit': 'Mpix/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': c['sample'], 'dwt_mpix_s': c['dwt_mpix_s'], 'dtcwt_mpix_s': c['dtcwt_mpix_s']}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'Mpix/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': 1e3 * t_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, world), 'parts': parts, 'roofline': roof, 'cpu_baseline': cpu,
            'e2e': e2e, 'clocks': clocks, 'gpu_launches': launches,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(torch, dist, pw, dev, world, dshape, tshape, steps):
    """Same step through the public nn.Module API, inputs in pinned host memory, every output copied back to
    pinned host memory, all inside the timed region.  The batch is processed in chunks on three streams
    (H2D / compute / D2H) so copies overlap compute; PCIe is the bound."""
    from pytorch_wavelets_b200 import pipeline
    dwt = pw.DWTForward(J=3, wave='db4', mode='symmetric').to(dev)
    dtc = pw.DTCWTForward(J=3).to(dev)
    hd = torch.empty(dshape, pin_memory=True)
    ht = torch.empty(tshape, pin_memory=True)
    hd[:8].normal_()
    ht[:8].normal_()
    for i in range(8, hd.shape[0], 8):      # synthetic host data: a few random images repeated (cheap to generate)
        hd[i:i + 8].copy_(hd[:min(8, hd.shape[0] - i)])
    for i in range(8, ht.shape[0], 8):
        ht[i:i + 8].copy_(ht[:min(8, ht.shape[0] - i)])
    pd = pipeline.HostPipeline(dwt, hd.shape, dev, chunk=16)
    pt = pipeline.HostPipeline(dtc, ht.shape, dev, chunk=8)
    for _ in range(2):
        pd.run(hd)
        pt.run(ht)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        pd.run(hd)
        pt.run(ht)
    t1.record()
    torch.cuda.synchronize()
    t = t0.elapsed_time(t1) / 1e3
    tt = torch.tensor([t], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
    pix = hd.numel() + ht.numel()
    return {'value': world * pix * steps / t / 1e6, 'unit': 'Mpix/s',
            'h2d_bytes_per_step': 4 * pix, 'd2h_bytes_per_step': pd.out_bytes + pt.out_bytes,
            'ms_per_step': 1e3 * t / steps, 'steps': steps,
            'how': 'pinned host in/out, chunked 3-stream pipeline through DWTForward/DTCWTForward.forward'}


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
