"""The reference's GPU op sequence for the two headline transforms, written independently from the closed forms of
SURVEY.md section 8 (extension by index gather, depthwise F.conv2d, slicing / stacking) -- the "existing ATen kernels"
bar that `bench.py --impl aten` times on the same B200.  NOT part of the product (the product never imports this);
tests/test_aten_arm.py checks on the CPU that it computes the same operator as the oracle.
All filter arguments are the module buffers (stored / reversed taps, 1-D tensors)."""
import numpy as np
import torch
import torch.nn.functional as F


def sym_idx(n, lo, hi, dev):   # half-sample symmetric extension of range(n) over [lo, hi)
    i = np.arange(lo, hi)
    r = np.mod(i, 2 * n)
    return torch.as_tensor(np.where(r < n, r, 2 * n - 1 - r), device=dev)

def analysis_1d(x, f_lo, f_hi, dim):   # out[k] = sum_j f[j] xe[2k + j - (L-2)], stored (reversed) taps, symmetric
    L = f_lo.numel()
    n = x.shape[dim]
    K = (n + L - 1) // 2
    xe = x.index_select(dim, sym_idx(n, -(L - 2), 2 * K, x.device))
    C = x.shape[1]
    w = torch.stack((f_lo, f_hi)).reshape(2, 1, -1)
    w = w.repeat(C, 1, 1)
    if dim == 3:
        return F.conv2d(xe, w.reshape(2 * C, 1, 1, L), stride=(1, 2), groups=C)
    return F.conv2d(xe, w.reshape(2 * C, 1, L, 1), stride=(2, 1), groups=C)

def dwt_fwd(x, f_lo, f_hi, J):
    yh = []
    ll = x
    for _ in range(J):
        N, C, H, W = ll.shape
        lohi = analysis_1d(ll, f_lo, f_hi, 3)                 # (N, 2C, H, W'): [lo_c, hi_c] interleaved
        y = analysis_1d(lohi, f_lo, f_hi, 2)                  # (N, 4C, H', W'): per input channel ll, lh, hl, hh
        y = y.reshape(N, C, 4, y.shape[-2], y.shape[-1])
        ll = y[:, :, 0].contiguous()
        yh.append(y[:, :, 1:].contiguous())
    return ll, yh

def filt_same(x, h, dim):   # undecimated odd-length filter with symmetric extension (level-1 DTCWT)
    L = h.numel()
    n = x.shape[dim]
    m = L // 2
    xe = x.index_select(dim, sym_idx(n, -m, n + m, x.device))
    C = x.shape[1]
    w = h.reshape(1, 1, -1).repeat(C, 1, 1)
    return F.conv2d(xe, w.reshape(C, 1, 1, L) if dim == 3 else w.reshape(C, 1, L, 1), groups=C)

def dfilt(x, ha, hb, highpass, dim):   # dual-tree decimation by 2, stored (reversed) taps
    m = ha.numel()
    n = x.shape[dim]
    C = x.shape[1]
    xe = x.index_select(dim, sym_idx(n, -(m - 2), n + m, x.device))
    def tree(h, off):
        xs = xe.narrow(dim, off, xe.shape[dim] - off)
        w = h.reshape(1, 1, -1).repeat(C, 1, 1)
        if dim == 3:
            return F.conv2d(xs, w.reshape(C, 1, 1, m), stride=(1, 4), dilation=(1, 2), groups=C)
        return F.conv2d(xs, w.reshape(C, 1, m, 1), stride=(4, 1), dilation=(2, 1), groups=C)
    ya, yb = tree(ha, 0), tree(hb, 1)
    q = n // 4
    ya, yb = ya.narrow(dim, 0, q), yb.narrow(dim, 0, q)
    pair = (yb, ya) if highpass else (ya, yb)
    out = torch.stack(pair, dim=dim + 1)
    shp = list(x.shape)
    shp[dim] = n // 2
    return out.reshape(shp)

def q2c(y):
    y = y * (0.5 ** 0.5)
    a, b, c, d = y[..., 0::2, 0::2], y[..., 0::2, 1::2], y[..., 1::2, 0::2], y[..., 1::2, 1::2]
    return (a - d, b + c), (a + d, b - c)

def pack(lh, hl, hh):
    (l1, l2), (v1, v2), (d1, d2) = q2c(lh), q2c(hl), q2c(hh)
    re = torch.stack((l1[0], d1[0], v1[0], v2[0], d2[0], l2[0]), dim=2)
    im = torch.stack((l1[1], d1[1], v1[1], v2[1], d2[1], l2[1]), dim=2)
    return torch.stack((re, im), dim=-1)

def dtcwt_fwd(x, h0o, h1o, h0a, h0b, h1a, h1b, J):
    lo, hi = filt_same(x, h0o, 3), filt_same(x, h1o, 3)
    ll = filt_same(lo, h0o, 2)
    yh = [pack(filt_same(lo, h1o, 2), filt_same(hi, h0o, 2), filt_same(hi, h1o, 2))]
    for _ in range(1, J):
        lo, hi = dfilt(ll, h0b, h0a, False, 3), dfilt(ll, h1b, h1a, True, 3)
        lh, hl, hh = dfilt(lo, h1b, h1a, True, 2), dfilt(hi, h0b, h0a, False, 2), dfilt(hi, h1b, h1a, True, 2)
        ll = dfilt(lo, h0b, h0a, False, 2)
        yh.append(pack(lh, hl, hh))
    return ll, yh

