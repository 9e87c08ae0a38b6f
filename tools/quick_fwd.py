import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi
x = torch.randn(128, 32, 512, 512, device='cuda'); f = pw.DWTForward(J=3, wave='db4', mode='symmetric').cuda()
xt = torch.randn(64, 3, 1024, 1024, device='cuda'); d = pw.DTCWTForward(J=3).cuda()
import ctypes
_ffi.lib().b200w_debug_set_want.argtypes = [ctypes.c_int]
xs = torch.randn(256, 3, 256, 256, device='cuda'); sc = torch.nn.Sequential(pw.ScatLayer(), pw.ScatLayer()).cuda()
for want in (0,):
  _ffi.lib().b200w_debug_set_want(want)
  res = {'want': want}
  with torch.no_grad():
    for name, fn, inp in (('dwt', f, x), ('dtcwt', d, xt), ('scat', sc, xs)):
        for _ in range(3): fn(inp)
        torch.cuda.synchronize()
        rec = _ffi.CallRecorder()
        with rec:
            for _ in range(10): fn(inp)
        s = rec.summary()
        res[name] = {' '.join(k.split()[:2]).replace('dtcwt_','').replace('dwt_',''): round(v['avg_ms'], 4) for k, v in sorted(s.items())}
        res[name]['total'] = round(sum(v['avg_ms'] for v in s.values()), 4)
  print(json.dumps(res))
