import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi
L = _ffi.lib()
L.b200w_debug_set_minb.argtypes = [ctypes.c_int]
x = torch.randn(128, 32, 512, 512, device='cuda')
f = pw.DWTForward(J=3, wave='db4', mode='symmetric').cuda()
L.b200w_debug_set_hs.argtypes = [ctypes.c_int]
for hs, minb in ((2, 0), (2, 1), (4, 0), (2, 0), (2, 1), (4, 0)):
    L.b200w_debug_set_minb(minb); L.b200w_debug_set_hs(hs)
    with torch.no_grad():
        for _ in range(3): f(x)
        torch.cuda.synchronize()
        rec = _ffi.CallRecorder()
        with rec:
            for _ in range(10): f(x)
        s = rec.summary()
    print('hs', hs, 'minb', minb, {k.split()[1]: round(v['avg_ms'], 4) for k, v in sorted(s.items())}, 'total', round(sum(v['avg_ms'] for v in s.values()), 4))
