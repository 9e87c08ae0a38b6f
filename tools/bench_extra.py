"""Secondary timings (not the bench.py line): inverses and ScatLayer x2 at the BASELINE.json shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw

dev = 'cuda'
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
with torch.no_grad():
    x = torch.randn(128, 32, 512, 512, device=dev)
    f = pw.DWTForward(J=3, wave='db4', mode='symmetric').to(dev); g = pw.DWTInverse(wave='db4', mode='symmetric').to(dev)
    c = f(x)
    out['dwt_fwd_ms'] = timeit(lambda: f(x)); out['dwt_inv_ms'] = timeit(lambda: g(c))
    out['dwt_pr_err'] = float((g(c) - x).abs().max())
    del x, c
    x = torch.randn(64, 3, 1024, 1024, device=dev)
    f = pw.DTCWTForward(J=3).to(dev); g = pw.DTCWTInverse().to(dev)
    c = f(x)
    out['dtcwt_fwd_ms'] = timeit(lambda: f(x)); out['dtcwt_inv_ms'] = timeit(lambda: g(c))
    out['dtcwt_pr_err'] = float((g(c) - x).abs().max())
    del x, c
    x = torch.randn(256, 3, 256, 256, device=dev)
    s = torch.nn.Sequential(pw.ScatLayer(), pw.ScatLayer()).to(dev)
    out['scat2_ms'] = timeit(lambda: s(x))
    out['dwt_fwd_gpix_s'] = 128*32*512*512 / out['dwt_fwd_ms'] / 1e6; out['dwt_inv_gpix_s'] = 128*32*512*512 / out['dwt_inv_ms'] / 1e6
    out['dtcwt_fwd_gpix_s'] = 64*3*1024*1024 / out['dtcwt_fwd_ms'] / 1e6; out['dtcwt_inv_gpix_s'] = 64*3*1024*1024 / out['dtcwt_inv_ms'] / 1e6
    out['scat2_gpix_s'] = 256*3*256*256 / out['scat2_ms'] / 1e6
    del x
    # BASELINE configs[4] per-GPU shard shape (reduced batch): DWT J=4 db8 zero on 2048x2048
    x = torch.randn(8, 16, 2048, 2048, device=dev)
    f = pw.DWTForward(J=4, wave='db8', mode='zero').to(dev); g = pw.DWTInverse(wave='db8', mode='zero').to(dev)
    c = f(x)
    out['c5_dwt_fwd_ms'] = timeit(lambda: f(x)); out['c5_dwt_inv_ms'] = timeit(lambda: g(c))
    out['c5_dwt_fwd_gpix_s'] = x.numel() / out['c5_dwt_fwd_ms'] / 1e6; out['c5_dwt_inv_gpix_s'] = x.numel() / out['c5_dwt_inv_ms'] / 1e6
print(json.dumps({k: (v if k.endswith('_err') else round(v, 4)) for k, v in out.items()}))
