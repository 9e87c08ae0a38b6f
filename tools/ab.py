"""A/B timing of the public transforms at the BASELINE shapes (CUDA events, 10 launches after 3 warm-ups).
Run under B200W_LIB=<variant .so> to time a variant build; prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw

which = set(sys.argv[1:]) or {'dwt', 'dwti', 'dtcwt', 'dtcwti', 'scat', 'c5'}


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4)


out = {'lib': os.environ.get('B200W_LIB', 'default')}
with torch.no_grad():
    if which & {'dwt', 'dwti'}:
        x = torch.randn(128, 32, 512, 512, device='cuda')
        for J in (1, 3):
            f = pw.DWTForward(J=J, wave='db4', mode='symmetric').cuda()
            if 'dwt' in which:
                out['dwt_J%d' % J] = timeit(lambda: f(x))
        if 'dwti' in which:
            yl, yh = f(x)
            g = pw.DWTInverse(wave='db4', mode='symmetric').cuda()
            out['dwt_inv'] = timeit(lambda: g((yl, yh)))
            del yl, yh
        del x
    if which & {'dtcwt', 'dtcwti'}:
        x = torch.randn(64, 3, 1024, 1024, device='cuda')
        f = pw.DTCWTForward(J=3).cuda()
        if 'dtcwt' in which:
            out['dtcwt_fwd'] = timeit(lambda: f(x))
        if 'dtcwti' in which:
            yl, yh = f(x)
            g = pw.DTCWTInverse().cuda()
            out['dtcwt_inv'] = timeit(lambda: g((yl, yh)))
            del yl, yh
        del x
    if 'scat' in which:
        from pytorch_wavelets_b200 import ScatLayer
        x = torch.randn(256, 3, 256, 256, device='cuda')
        s = torch.nn.Sequential(ScatLayer(), ScatLayer()).cuda()
        out['scat2'] = timeit(lambda: s(x))
        s1 = ScatLayer().cuda()
        out['scat_l1'] = timeit(lambda: s1(x))
        del x
    if 'c5' in which:
        x = torch.randn(8, 16, 2048, 2048, device='cuda')
        f = pw.DWTForward(J=4, wave='db8', mode='zero').cuda()
        out['c5_chunk'] = timeit(lambda: f(x))
        del x
print(json.dumps(out))
