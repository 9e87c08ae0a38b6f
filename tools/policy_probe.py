"""Policy probe: whole-transform time of DWTForward(J=3, db4, symmetric) for several plane sizes at a constant pixel count.
Run once with the shipped library and once with B200W_LIB=<variant built with -DB200W_PYR_FUSE_ALL_MAX_SIDE=100000>."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
out = {}
with torch.no_grad():
    for side in (64, 128, 256, 512, 1024):
        planes = (1 << 28) // (side * side)          # 268 Mpix per call
        x = torch.randn(planes // 8, 8, side, side, device='cuda')
        for J in (2, 3):
            f = pw.DWTForward(J=J, wave='db4', mode='symmetric').cuda()
            for _ in range(3): f(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): f(x)
            e1.record(); torch.cuda.synchronize()
            out['%d_J%d' % (side, J)] = round(e0.elapsed_time(e1) / 10, 4)
        del x
print(json.dumps(out))
