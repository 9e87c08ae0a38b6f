"""numpy front-end of the CPU oracle (oracle/wave_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  Nothing under pytorch_wavelets_b200/ imports it.

Two layers:
  * thin per-level wrappers with the argument meaning of include/b200wave.h (host numpy arrays);
  * module-level compositions restating the reference's nn.Module.forward loops:
      dwt_forward / dwt_inverse        dwt/transform2d.py:44-74, 111-148
      dtcwt_forward / dtcwt_inverse    dtcwt/transform2d.py:87-147, 193-254
      scat_layer                       scatternet/layers.py:51-75
    (paths relative to /root/reference/pytorch_wavelets/).
Filters are given in the reference's *stored* form (what its module buffers hold).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODES = {'zero': 0, 'symmetric': 1, 'per': 2, 'periodization': 2, 'constant': 3, 'reflect': 4,
         'replicate': 5, 'periodic': 6}


def build(force=False):
    so = os.path.join(_HERE, 'libwaveoracle.so')
    srcs = [os.path.join(_HERE, f) for f in ('wave_oracle.c', 'wave_oracle_impl.h', 'Makefile')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, '-s', 'CC=gcc'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'libwaveoracle.so')
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return '_f32', ctypes.c_float
    if dtype == np.float64:
        return '_f64', ctypes.c_double
    raise TypeError(dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _taps(f, dtype):
    return np.ascontiguousarray(np.asarray(f, dtype=np.float64).ravel().astype(dtype))


def _check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed with code %d' % (what, rc))


def mode_int(mode):
    if isinstance(mode, str):
        if mode not in MODES:
            raise ValueError("Unkown pad type: {}".format(mode))
        return MODES[mode]
    return int(mode)


def coeff_len(n, flen, mode):
    return lib().orc_coeff_len(int(n), int(flen), mode_int(mode))


def rec_len(k, flen, mode):
    return lib().orc_rec_len(int(k), int(flen), mode_int(mode))


# ------------------------------------------------------------------------------------------------
# per-level entry points (same meaning as include/b200wave.h)

def dwt_afb2d(x, fw_lo, fw_hi, fh_lo, fh_hi, mode):
    """x (N,C,H,W) -> ll (N,C,Ho,Wo), highs (N,C,3,Ho,Wo)."""
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    m = mode_int(mode)
    fw_lo, fw_hi, fh_lo, fh_hi = [_taps(f, x.dtype) for f in (fw_lo, fw_hi, fh_lo, fh_hi)]
    Lw, Lh = fw_lo.size, fh_lo.size
    Ho, Wo = coeff_len(H, Lh, m), coeff_len(W, Lw, m)
    ll = np.empty((N, C, Ho, Wo), x.dtype)
    highs = np.empty((N, C, 3, Ho, Wo), x.dtype)
    fn = getattr(lib(), 'orc_dwt_afb2d' + sfx)
    rc = fn(_p(x), ctypes.c_longlong(H * W), W, _p(ll), ctypes.c_longlong(Ho * Wo), Wo, _p(highs),
            N * C, H, W, _p(fw_lo), _p(fw_hi), Lw, _p(fh_lo), _p(fh_hi), Lh, m)
    if rc == -1:
        raise ValueError("Unkown pad type: {}".format(mode))
    _check(rc, 'orc_dwt_afb2d')
    return ll, highs


def dwt_sfb2d(ll, highs, gh_lo, gh_hi, gw_lo, gw_hi, mode, out_hw=None):
    """ll (N,C,Hc,Wc), highs (N,C,3,Hc,Wc) or None -> y (N,C,Ho,Wo)."""
    ll = np.ascontiguousarray(ll)
    sfx, _ = _sfx(ll.dtype)
    N, C, Hc, Wc = ll.shape
    m = mode_int(mode)
    if highs is not None:
        highs = np.ascontiguousarray(highs, dtype=ll.dtype)
        assert highs.shape == (N, C, 3, Hc, Wc), (highs.shape, ll.shape)
    gh_lo, gh_hi, gw_lo, gw_hi = [_taps(f, ll.dtype) for f in (gh_lo, gh_hi, gw_lo, gw_hi)]
    Lh, Lw = gh_lo.size, gw_lo.size
    Ho, Wo = rec_len(Hc, Lh, m), rec_len(Wc, Lw, m)
    if out_hw is not None:
        Ho, Wo = min(Ho, out_hw[0]), min(Wo, out_hw[1])
    y = np.empty((N, C, Ho, Wo), ll.dtype)
    fn = getattr(lib(), 'orc_dwt_sfb2d' + sfx)
    rc = fn(_p(ll), ctypes.c_longlong(Hc * Wc), Wc, _p(highs), _p(y), ctypes.c_longlong(Ho * Wo), Wo,
            N * C, Hc, Wc, Ho, Wo, _p(gh_lo), _p(gh_hi), Lh, _p(gw_lo), _p(gw_hi), Lw, m)
    if rc == -1:
        raise ValueError("Unkown pad type: {}".format(mode))
    _check(rc, 'orc_dwt_sfb2d')
    return y


def dwt_afb1d(x, f0, f1, mode):
    """AFB1D.forward (dwt/lowlevel.py:388-404): x (N,C,L) -> lo, hi (N,C,K)."""
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    N, C, n = x.shape
    m = mode_int(mode)
    f0, f1 = _taps(f0, x.dtype), _taps(f1, x.dtype)
    K = coeff_len(n, f0.size, m)
    lo, hi = np.empty((N, C, K), x.dtype), np.empty((N, C, K), x.dtype)
    rc = getattr(lib(), 'orc_dwt_afb1d' + sfx)(_p(x), ctypes.c_longlong(n), N * C, n, _p(lo), _p(hi), _p(f0), _p(f1),
                                               f0.size, m)
    _check(rc, 'orc_dwt_afb1d')
    return lo, hi


def dwt_sfb1d(lo, hi, g0, g1, mode, out_len=None):
    """SFB1D.forward (dwt/lowlevel.py:717-729): lo, hi (N,C,K) -> y (N,C,rec_len); hi may be None (zeros)."""
    lo = np.ascontiguousarray(lo)
    sfx, _ = _sfx(lo.dtype)
    N, C, K = lo.shape
    m = mode_int(mode)
    g0, g1 = _taps(g0, lo.dtype), _taps(g1, lo.dtype)
    n = rec_len(K, g0.size, m) if out_len is None else int(out_len)
    if hi is not None:
        hi = np.ascontiguousarray(hi)
    y = np.empty((N, C, n), lo.dtype)
    rc = getattr(lib(), 'orc_dwt_sfb1d' + sfx)(_p(lo), _p(hi), N * C, K, _p(y), n, _p(g0), _p(g1), g0.size, m)
    _check(rc, 'orc_dwt_sfb1d')
    return y


def dwt1d_forward(x, filts, J, mode):
    """DWT1DForward.forward (dwt/transform1d.py:44-65); filts = stored (h0, h1)."""
    x0, highs = x, []
    for _ in range(J):
        x0, x1 = dwt_afb1d(x0, filts[0], filts[1], mode)
        highs.append(x1)
    return x0, highs


def dwt1d_inverse(yl, yh, filts, mode):
    """DWT1DInverse.forward (dwt/transform1d.py:97-115): None band-passes are zeros; the low-pass loses its last
    sample when it is one longer than the band-pass ('unpad')."""
    x0 = yl
    for x1 in yh[::-1]:
        if x1 is not None and x0.shape[-1] > x1.shape[-1]:
            x0 = x0[..., :-1]
        x0 = dwt_sfb1d(x0, x1, filts[0], filts[1], mode)
    return x0


def highs_shape_strides(N, C, h, w, o_dim=2, ri_dim=-1):
    """Shape of the reference's 6-D band-pass tensor and its element strides in the order
    (n, c, orientation, row, col, real/imag).  Restates get_dimensions6, dtcwt/transform_funcs.py:32-58
    (the stack at o_dim of the 4-D bands, then the stack at ri_dim of the 5-D result)."""
    o5 = o_dim % 6
    ri = ri_dim % 6
    if ri < o5:
        o5 -= 1
    dims5 = ['n', 'c', 'h', 'w']
    dims5.insert(o5, 'o')
    dims6 = list(dims5)
    dims6.insert(ri, 'r')
    size = {'n': N, 'c': C, 'o': 6, 'h': h, 'w': w, 'r': 2}
    shape = tuple(size[d] for d in dims6)
    strides = {}
    acc = 1
    for d in reversed(dims6):
        strides[d] = acc
        acc *= size[d]
    hs = [strides[k] for k in ('n', 'c', 'o', 'h', 'w', 'r')]
    return shape, hs


def _hs(hs):
    return (ctypes.c_longlong * 6)(*[int(v) for v in hs])


def dtcwt_fwd_j1(x, h0, h1, skip_hps=False, o_dim=2, ri_dim=-1, mode='symmetric'):
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    h0, h1 = _taps(h0, x.dtype), _taps(h1, x.dtype)
    ll = np.empty((N, C, H, W), x.dtype)
    shape, hs = highs_shape_strides(N, C, H // 2, W // 2, o_dim, ri_dim)
    highs = None if skip_hps else np.empty(shape, x.dtype)
    fn = getattr(lib(), 'orc_dtcwt_fwd_j1' + sfx)
    rc = fn(_p(x), ctypes.c_longlong(H * W), W, _p(ll), ctypes.c_longlong(H * W), W, _p(highs), _hs(hs),
            N, C, H, W, _p(h0), h0.size, _p(h1), h1.size, mode_int(mode))
    _check(rc, 'orc_dtcwt_fwd_j1')
    return ll, highs


def dtcwt_fwd_j2plus(x, h0a, h1a, h0b, h1b, skip_hps=False, o_dim=2, ri_dim=-1):
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    if H % 4 != 0:
        raise ValueError('No. of rows in X must be a multiple of 4\nX was {}'.format(x.shape))
    if W % 4 != 0:
        raise ValueError('No. of cols in X must be a multiple of 4\nX was {}'.format(x.shape))
    h0a, h1a, h0b, h1b = [_taps(f, x.dtype) for f in (h0a, h1a, h0b, h1b)]
    ll = np.empty((N, C, H // 2, W // 2), x.dtype)
    shape, hs = highs_shape_strides(N, C, H // 4, W // 4, o_dim, ri_dim)
    highs = None if skip_hps else np.empty(shape, x.dtype)
    fn = getattr(lib(), 'orc_dtcwt_fwd_j2plus' + sfx)
    rc = fn(_p(x), ctypes.c_longlong(H * W), W, _p(ll), ctypes.c_longlong((H // 2) * (W // 2)), W // 2,
            _p(highs), _hs(hs), N, C, H, W, _p(h0a), _p(h1a), _p(h0b), _p(h1b), h0a.size)
    _check(rc, 'orc_dtcwt_fwd_j2plus')
    return ll, highs


def dtcwt_inv_j1(ll, highs, g0, g1, o_dim=2, ri_dim=-1, mode='symmetric'):
    """ll (N,C,H,W) or None, highs 6-D or None -> y (N,C,H,W)."""
    ref = ll if ll is not None else highs
    dtype = ref.dtype
    sfx, _ = _sfx(dtype)
    if highs is not None:
        highs = np.ascontiguousarray(highs)
        names = _dim_names(o_dim, ri_dim)
        sz = dict(zip(names, highs.shape))
        if ll is not None:
            # inv_j1 itself cuts ll back to match the highs, transform_funcs.py:170-176
            if ll.shape[2] != 2 * sz['h']:
                ll = ll[:, :, 1:-1]
            if ll.shape[3] != 2 * sz['w']:
                ll = ll[:, :, :, 1:-1]
    if ll is not None:
        ll = np.ascontiguousarray(ll)
        N, C, H, W = ll.shape
    if highs is not None:
        if ll is None:
            N, C, H, W = sz['n'], sz['c'], 2 * sz['h'], 2 * sz['w']
        _, hs = highs_shape_strides(N, C, H // 2, W // 2, o_dim, ri_dim)
    else:
        hs = [0] * 6
        mode = 'symmetric'   # reference quirk: the low-pass-only path ignores `mode` (transform_funcs.py:158-159)
    g0, g1 = _taps(g0, dtype), _taps(g1, dtype)
    y = np.empty((N, C, H, W), dtype)
    fn = getattr(lib(), 'orc_dtcwt_inv_j1' + sfx)
    rc = fn(_p(ll), ctypes.c_longlong(H * W), W, _p(highs), _hs(hs), _p(y), ctypes.c_longlong(H * W), W,
            N, C, H, W, _p(g0), g0.size, _p(g1), g1.size, mode_int(mode))
    _check(rc, 'orc_dtcwt_inv_j1')
    return y


def dtcwt_inv_j2plus(ll, highs, g0a, g1a, g0b, g1b, o_dim=2, ri_dim=-1):
    """ll (N,C,H,W) or None, highs 6-D at (H/2,W/2) or None -> y (N,C,2H,2W)."""
    ref = ll if ll is not None else highs
    dtype = ref.dtype
    sfx, _ = _sfx(dtype)
    if ll is not None:
        ll = np.ascontiguousarray(ll)
        N, C, H, W = ll.shape
    if highs is not None:
        highs = np.ascontiguousarray(highs)
        names = _dim_names(o_dim, ri_dim)
        sz = dict(zip(names, highs.shape))
        if ll is None:
            N, C, H, W = sz['n'], sz['c'], 2 * sz['h'], 2 * sz['w']
        _, hs = highs_shape_strides(N, C, H // 2, W // 2, o_dim, ri_dim)
    else:
        hs = [0] * 6
    g0a, g1a, g0b, g1b = [_taps(f, dtype) for f in (g0a, g1a, g0b, g1b)]
    y = np.empty((N, C, 2 * H, 2 * W), dtype)
    fn = getattr(lib(), 'orc_dtcwt_inv_j2plus' + sfx)
    rc = fn(_p(ll), ctypes.c_longlong(H * W), W, _p(highs), _hs(hs), _p(y), ctypes.c_longlong(4 * H * W), 2 * W,
            N, C, H, W, _p(g0a), _p(g1a), _p(g0b), _p(g1b), g0a.size)
    _check(rc, 'orc_dtcwt_inv_j2plus')
    return y


def _dim_names(o_dim, ri_dim):
    o5 = o_dim % 6
    ri = ri_dim % 6
    if ri < o5:
        o5 -= 1
    d = ['n', 'c', 'h', 'w']
    d.insert(o5, 'o')
    d.insert(ri, 'r')
    return d


def scat_j1(x, h0, h1, mode='symmetric', magbias=1e-2, want_grad_aux=False):
    x = np.ascontiguousarray(x)
    sfx, ct = _sfx(x.dtype)
    N, C, H, W = x.shape
    h0, h1 = _taps(h0, x.dtype), _taps(h1, x.dtype)
    z = np.empty((N, 7, C, H // 2, W // 2), x.dtype)
    dre = dim = None
    if want_grad_aux:
        dre = np.empty((N, 6, C, H // 2, W // 2), x.dtype)
        dim = np.empty_like(dre)
    fn = getattr(lib(), 'orc_scat_j1' + sfx)
    rc = fn(_p(x), _p(z), _p(dre), _p(dim), N, C, H, W, _p(h0), h0.size, _p(h1), h1.size, mode_int(mode),
            ct(magbias))
    _check(rc, 'orc_scat_j1')
    return (z, dre, dim) if want_grad_aux else z


# 1-D primitives (unit tests against the reference's colfilter/coldfilt/colifilt)

def filter1d(x, h, symmetric=True, along_w=False):
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    h = _taps(h, x.dtype)
    ext = 2 * (h.size // 2) - h.size + 1
    y = np.empty((N, C, H + (0 if along_w else ext), W + (ext if along_w else 0)), x.dtype)
    _check(getattr(lib(), 'orc_filter' + sfx)(_p(x), _p(y), N * C, H, W, _p(h), h.size, int(symmetric),
                                              int(along_w)), 'orc_filter')
    return y


def dfilt1d(x, ha, hb, highpass=False, along_w=False):
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    ha, hb = _taps(ha, x.dtype), _taps(hb, x.dtype)
    y = np.empty((N, C, H if along_w else H // 2, W // 2 if along_w else W), x.dtype)
    rc = getattr(lib(), 'orc_dfilt' + sfx)(_p(x), _p(y), N * C, H, W, _p(ha), _p(hb), ha.size, int(highpass),
                                           int(along_w))
    if rc == -2:
        raise ValueError('No. of %s in X must be a multiple of 4' % ('cols' if along_w else 'rows'))
    _check(rc, 'orc_dfilt')
    return y


def ifilt1d(x, ha, hb, highpass=False, along_w=False):
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    ha, hb = _taps(ha, x.dtype), _taps(hb, x.dtype)
    y = np.empty((N, C, H if along_w else 2 * H, 2 * W if along_w else W), x.dtype)
    rc = getattr(lib(), 'orc_ifilt' + sfx)(_p(x), _p(y), N * C, H, W, _p(ha), _p(hb), ha.size, int(highpass),
                                           int(along_w))
    if rc == -2:
        raise ValueError('No. of %s in X must be a multiple of 2' % ('cols' if along_w else 'rows'))
    _check(rc, 'orc_ifilt')
    return y


# ------------------------------------------------------------------------------------------------
# module-level compositions

def dwt_forward(x, filts, J, mode):
    """DWTForward.forward, dwt/transform2d.py:63-74.  filts = stored (h0_col,h1_col,h0_row,h1_row);
    the *_col pair is applied along W and the *_row pair along H (the argument-order quirk at
    transform2d.py:70-71 vs lowlevel.py:336)."""
    h0_col, h1_col, h0_row, h1_row = filts
    yh = []
    ll = x
    for _ in range(J):
        ll, high = dwt_afb2d(ll, h0_col, h1_col, h0_row, h1_row, mode)
        yh.append(high)
    return ll, yh


def dwt_inverse(yl, yh, filts, mode):
    """DWTInverse.forward, dwt/transform2d.py:131-148.  filts = stored (g0_col,g1_col,g0_row,g1_row);
    the *_col pair ends up along W and the *_row pair along H (same swap, :146-147 vs lowlevel.py:671)."""
    g0_col, g1_col, g0_row, g1_row = filts
    ll = yl
    for h in yh[::-1]:
        if h is None:
            h = np.zeros((ll.shape[0], ll.shape[1], 3, ll.shape[-2], ll.shape[-1]), ll.dtype)
        if ll.shape[-2] > h.shape[-2]:
            ll = ll[..., :-1, :]
        if ll.shape[-1] > h.shape[-1]:
            ll = ll[..., :-1]
        # SFB2D.forward(low, highs, g0_row:=g0_col, g1_row:=g1_col, g0_col:=g0_row, g1_col:=g1_row):
        # first pass (dim=2, along H) uses the module's *_row buffers, second (along W) the *_col ones.
        ll = dwt_sfb2d(ll, h, g0_row, g1_row, g0_col, g1_col, mode)
    return ll


def dtcwt_forward(x, level1, qshift, J=3, skip_hps=False, include_scale=False, o_dim=2, ri_dim=-1,
                  mode='symmetric'):
    """DTCWTForward.forward, dtcwt/transform2d.py:108-147.  level1 = stored (h0o, h1o);
    qshift = stored (h0a, h0b, h1a, h1b).  Skipped band-passes are None here (0-dim tensors there)."""
    h0o, h1o = level1
    h0a, h0b, h1a, h1b = qshift
    if not isinstance(skip_hps, (list, tuple)):
        skip_hps = [skip_hps] * J
    if not isinstance(include_scale, (list, tuple)):
        include_scale = [include_scale] * J
    scales = [None] * J
    highs = [None] * J
    if J == 0:
        return x, None
    r, c = x.shape[2:]
    if r % 2 != 0:
        x = np.concatenate((x, x[:, :, -1:]), axis=2)
    if c % 2 != 0:
        x = np.concatenate((x, x[:, :, :, -1:]), axis=3)
    low, h = dtcwt_fwd_j1(x, h0o, h1o, skip_hps[0], o_dim, ri_dim, mode)
    highs[0] = h
    if include_scale[0]:
        scales[0] = low
    for j in range(1, J):
        r, c = low.shape[2:]
        if r % 4 != 0:
            low = np.concatenate((low[:, :, 0:1], low, low[:, :, -1:]), axis=2)
        if c % 4 != 0:
            low = np.concatenate((low[:, :, :, 0:1], low, low[:, :, :, -1:]), axis=3)
        low, h = dtcwt_fwd_j2plus(low, h0a, h1a, h0b, h1b, skip_hps[j], o_dim, ri_dim)
        highs[j] = h
        if include_scale[j]:
            scales[j] = low
    if True in include_scale:
        return scales, highs
    return low, highs


def dtcwt_inverse(yl, yh, level1, qshift, o_dim=2, ri_dim=-1, mode='symmetric'):
    """DTCWTInverse.forward, dtcwt/transform2d.py:219-254.  level1 = stored (g0o, g1o);
    qshift = stored (g0a, g0b, g1a, g1b)."""
    g0o, g1o = level1
    g0a, g0b, g1a, g1b = qshift
    low = yl
    J = len(yh)
    names = _dim_names(o_dim, ri_dim)
    h_dim, w_dim = names.index('h'), names.index('w')
    for s in yh[1:][::-1]:
        if s is not None:
            assert s.shape[o_dim] == 6, "Inverse transform must have input with 6 orientations"
            assert len(s.shape) == 6, "Bandpass inputs must have 6 dimensions"
            assert s.shape[ri_dim] == 2, "Inputs must be complex with real and imaginary parts in the ri dimension"
            r, c = low.shape[2:]
            r1, c1 = s.shape[h_dim], s.shape[w_dim]
            if r != r1 * 2:
                low = low[:, :, 1:-1]
            if c != c1 * 2:
                low = low[:, :, :, 1:-1]
        low = dtcwt_inv_j2plus(low, s, g0a, g1a, g0b, g1b, o_dim, ri_dim)
    if yh[0] is not None:
        r, c = low.shape[2:]
        r1, c1 = yh[0].shape[h_dim], yh[0].shape[w_dim]
        if r != r1 * 2:
            low = low[:, :, 1:-1]
        if c != c1 * 2:
            low = low[:, :, :, 1:-1]
    return dtcwt_inv_j1(low, yh[0], g0o, g1o, o_dim, ri_dim, mode)


def scat_layer(x, level1, mode='symmetric', magbias=1e-2):
    """ScatLayer.forward (combine_colour=False), scatternet/layers.py:51-75."""
    h0o, h1o = level1
    _, ch, r, c = x.shape
    if r % 2 != 0:
        x = np.concatenate((x, x[:, :, -1:]), axis=2)
    if c % 2 != 0:
        x = np.concatenate((x, x[:, :, :, -1:]), axis=3)
    z = scat_j1(x, h0o, h1o, mode, magbias)
    b, _, c_, h, w = z.shape
    return z.reshape(b, 7 * c_, h, w)
