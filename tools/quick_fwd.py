import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi
x = torch.randn(128, 32, 512, 512, device='cuda'); f = pw.DWTForward(J=3, wave='db4', mode='symmetric').cuda()
xt = torch.randn(64, 3, 1024, 1024, device='cuda'); d = pw.DTCWTForward(J=3).cuda()
res = {}
with torch.no_grad():
    for name, fn, inp in (('dwt', f, x), ('dtcwt', d, xt)):
        for _ in range(3): fn(inp)
        torch.cuda.synchronize()
        rec = _ffi.CallRecorder()
        with rec:
            for _ in range(10): fn(inp)
        s = rec.summary()
        res[name] = {k.split()[1]: round(v['avg_ms'], 4) for k, v in sorted(s.items())}
        res[name]['total'] = round(sum(v['avg_ms'] for v in s.values()), 4)
print(json.dumps(res))
