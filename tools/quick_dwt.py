import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi
x = torch.randn(128, 32, 512, 512, device='cuda'); f = pw.DWTForward(J=3, wave='db4', mode='symmetric').cuda()
with torch.no_grad():
    for _ in range(3): f(x)
    torch.cuda.synchronize()
    rec = _ffi.CallRecorder()
    with rec:
        for _ in range(10): f(x)
    s = rec.summary()
print(os.path.basename(_ffi.SO_PATH), {k.split()[1]: round(v['avg_ms'], 4) for k, v in sorted(s.items())})
