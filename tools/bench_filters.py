"""DTCWT forward / inverse with the other filter sets (64x3x1024x1024, J=3): streaming kernels vs generic tile kernels."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

lib = _ffi.lib()
x = torch.randn(64, 3, 1024, 1024, device='cuda')
out = {}
with torch.no_grad():
    for biort, qshift in (('near_sym_a', 'qshift_a'), ('near_sym_b', 'qshift_b'), ('antonini', 'qshift_c'), ('legall', 'qshift_d')):
        f = pw.DTCWTForward(biort=biort, qshift=qshift, J=3).cuda()
        g = pw.DTCWTInverse(biort=biort, qshift=qshift).cuda()
        c = f(x)
        r = {}
        r['stream_fwd_ms'] = round(timeit(lambda: f(x)), 3)
        r['stream_inv_ms'] = round(timeit(lambda: g(c)), 3)
        with _ffi.generic_kernels():
            r['generic_fwd_ms'] = round(timeit(lambda: f(x)), 3)
            r['generic_inv_ms'] = round(timeit(lambda: g(c)), 3)
        out[biort + '/' + qshift] = r
print(json.dumps(out))
