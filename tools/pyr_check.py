"""GPU check + timing of the fused DWT pyramid kernel against the level-by-level path (same arithmetic: bit-identical)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200.dwt import lowlevel
from pytorch_wavelets_b200 import _ffi

dev = 'cuda'


def per_level(f, x, J, mode):
    m = lowlevel.mode_to_int(mode)
    ll, yh = x, []
    for j in range(J):
        ll, h = lowlevel.afb2d_level(ll, f.h0_col, f.h1_col, f.h0_row, f.h1_row, m, pad_ll=(j < J - 1))
        yh.append(h)
    return ll.contiguous(), yh


def check(shape, J, wave, mode):
    torch.manual_seed(1)
    x = torch.randn(*shape, device=dev)
    f = pw.DWTForward(J=J, wave=wave, mode=mode).to(dev)
    L = f.h0_col.numel()
    N, C, H, W = shape
    wsb = _ffi.lib().b200w_dwt_forward_workspace(x.data_ptr(), H * W, W, N * C, H, W, J, L, L, lowlevel.mode_to_int(mode))
    print('case', shape, J, wave, mode, 'fused' if wsb == 0 else 'levels(ws=%d)' % wsb, flush=True)
    yl, yh = f(x)
    torch.cuda.synchronize()
    rl, rh = per_level(f, x, J, mode)
    ok = torch.equal(yl, rl) and all(torch.equal(a, b) for a, b in zip(yh, rh))
    if not ok:
        errs = [float((yl - rl).abs().max())] + [float((a - b).abs().max()) for a, b in zip(yh, rh)]
        nan = [bool(torch.isnan(yl).any())] + [bool(torch.isnan(a).any()) for a in yh]
        print('   MISMATCH max abs errs (yl, yh...):', errs, 'nan:', nan, flush=True)
        for j, (a, b) in enumerate(zip(yh, rh)):
            d = (a != b)
            if d.any():
                idx = d.nonzero()
                print('   yh%d: %d diffs; first' % (j, int(d.sum())), idx[0].tolist(), 'last', idx[-1].tolist(), flush=True)
        d = (yl != rl)
        if d.any():
            idx = d.nonzero()
            print('   yl: %d diffs; first' % int(d.sum()), idx[0].tolist(), 'last', idx[-1].tolist(), flush=True)
    return ok


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
    return e0.elapsed_time(e1) / reps


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    allok = True
    with torch.no_grad():
        if what in ('all', 'check'):
            cases = [((2, 3, 64, 64), 3, 'db4', 'symmetric'), ((2, 3, 64, 64), 1, 'db4', 'zero'),
                     ((1, 2, 96, 128), 3, 'db4', 'reflect'), ((2, 2, 99, 100), 2, 'db2', 'symmetric'),
                     ((1, 3, 130, 260), 4, 'db1', 'zero'), ((2, 2, 200, 256), 3, 'db3', 'symmetric'),
                     ((1, 2, 128, 128), 2, 'db5', 'symmetric'), ((1, 2, 160, 192), 2, 'db6', 'reflect'),
                     ((1, 2, 256, 256), 3, 'db8', 'zero'), ((3, 5, 512, 512), 3, 'db4', 'symmetric'),
                     ((1, 1, 1024, 1024), 3, 'db4', 'symmetric'), ((1, 2, 37, 52), 2, 'db4', 'zero')]
            for c in cases:
                ok = check(*c)
                print('   ->', 'OK' if ok else 'FAIL', flush=True)
                allok = allok and ok
        if what in ('all', 'time'):
            x = torch.randn(128, 32, 512, 512, device=dev)
            f = pw.DWTForward(J=3, wave='db4', mode='symmetric').to(dev)
            t_f = timeit(lambda: f(x))
            t_l = timeit(lambda: per_level(f, x, 3, 'symmetric'))
            alg = 4.0 * 4096 * (512 * 512 + 3 * (259 ** 2 + 133 ** 2 + 70 ** 2) + 70 ** 2)
            f1 = pw.DWTForward(J=1, wave='db4', mode='symmetric').to(dev)
            t_1 = timeit(lambda: f1(x))
            t_1l = timeit(lambda: per_level(f1, x, 1, 'symmetric'))
            alg1 = 4.0 * 4096 * (512 * 512 + 4 * 259 ** 2)
            print(json.dumps({'fused_J3_ms': t_f, 'levels_J3_ms': t_l, 'fused_J3_GBps': alg / t_f / 1e6,
                              'levels_J3_GBps': alg / t_l / 1e6, 'fused_J1_ms': t_1, 'levels_J1_ms': t_1l,
                              'fused_J1_GBps': alg1 / t_1 / 1e6, 'levels_J1_GBps': alg1 / t_1l / 1e6}), flush=True)
    print('ALL OK' if allok else 'SOME FAILED')
