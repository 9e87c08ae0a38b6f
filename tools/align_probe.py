"""Does the output row alignment (Wo odd / even / multiple of 8) change the DWT level-1 store throughput?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200.dwt import lowlevel as ll

big = torch.randn(128, 32, 512, 512, device='cuda')
w = pw.Wavelet('db4') if hasattr(pw, 'Wavelet') else None
f = pw.DWTForward(J=1, wave='db4', mode='symmetric').cuda()
taps = [f.h0_col, f.h1_col, f.h0_row, f.h1_row]
mode = ll.mode_to_int('symmetric')
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
import ctypes
from pytorch_wavelets_b200 import _ffi
_ffi.lib().b200w_debug_set_balanced.argtypes = [ctypes.c_int]
for balanced in [int(v) for v in os.environ.get('BALANCED', '0').split(',')]:
  _ffi.lib().b200w_debug_set_balanced(balanced)
  for S in [int(v) for v in os.environ.get('SIZES', '506,508,510,512,259,133').split(',')]:
    x = big[..., :S, :S] if S > 300 else big.view(-1, 32, 512, 512)[:128, :, :S, :S]
    lo, hi = ll.afb2d_level(x, *taps, mode)
    ms = t(lambda: ll.afb2d_level(x, *taps, mode))
    Ho, Wo = hi.shape[-2:]
    gb = 4 * 4096 * (S * S + 4 * Ho * Wo) / 1e9
    print(json.dumps({'balanced': balanced, 'S': S, 'Wo': Wo, 'ms': round(ms, 4), 'GBps': round(gb / ms * 1e3, 1)}))
