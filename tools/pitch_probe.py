"""Experiment: the per-warp DWT kernel writing the band-pass planes with a line-aligned row pitch (not the
reference layout) vs the contiguous layout, same shape, same bytes."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi
from pytorch_wavelets_b200.dwt import lowlevel as ll

lib = _ffi.lib()
lib.b200w_debug_set_hipitch.argtypes = [ctypes.c_int]
f = pw.DWTForward(J=1, wave='db4', mode='symmetric').cuda()
taps = [_ffi.host_taps(t) for t in (f.h0_col, f.h1_col, f.h0_row, f.h1_row)]
mode = ll.mode_to_int('symmetric')
P, S = 4096, 512
x = torch.randn(P, S, S, device='cuda')
Ho = Wo = (S + 7) // 2
def run(hip, llp):
    highs = torch.empty(P * 3 * Ho * hip + 64, device='cuda')
    low = torch.empty(P * Ho * llp + 64, device='cuda')
    def call():
        rc = lib.b200w_dwt_afb2d(x.data_ptr(), S * S, S, low.data_ptr(), Ho * llp, llp, highs.data_ptr(), P, S, S,
                                 taps[0].ptr, taps[1].ptr, taps[0].n, taps[2].ptr, taps[3].ptr, taps[2].n, mode,
                                 _ffi.stream_of(x))
        assert rc == 0
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
gb = 4 * P * (S * S + 4 * Ho * Wo) / 1e9
for hip, llp in ((Wo, Wo), (Wo, 288), (288, 288), (264, 264), (260, 260)):
    lib.b200w_debug_set_hipitch(0 if hip == Wo else hip)
    ms = run(hip, llp)
    print(json.dumps({'hipitch': hip, 'llpitch': llp, 'ms': round(ms, 4), 'GBps': round(gb / ms * 1e3, 1)}))
