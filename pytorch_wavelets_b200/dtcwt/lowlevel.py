"""Low-level DTCWT functions (reference ``pytorch_wavelets/dtcwt/lowlevel.py``).

In the transforms the 1-D passes are fused inside the per-level CUDA kernels (``transform_funcs.py``); the standalone
functions below exist for callers of the reference's low-level API -- same names, arguments, output sizes and
exceptions -- each one launch of a small CUDA kernel through the C ABI (``csrc/k_prims.cu``).  They are forward-only
(no autograd graph): the differentiable path is the ``FWD_J1 / FWD_J2PLUS / INV_J1 / INV_J2PLUS`` Functions.
``q2c`` / ``c2q`` are the reference's pointwise quad <-> complex re-packing (slicing and adds; also fused in the kernels).
"""
import numpy as np
import torch

from pytorch_wavelets_b200 import _ffi


def prep_filt(h, c, transpose=False):
    """Column filter -> reversed (c,1,L,1) tensor (or (c,1,1,L) with transpose), default dtype
    (reference dtcwt/lowlevel.py:58-67)."""
    h = np.atleast_2d(np.asarray(h))
    if h.shape[0] == 1:
        h = h.T
    h = h[::-1]
    h = h[None, None, :]
    h = np.repeat(h, repeats=c, axis=0)
    if transpose:
        h = h.transpose((0, 1, 3, 2))
    h = np.copy(h)
    return torch.tensor(h, dtype=torch.get_default_dtype())


def _is_nothing(X):
    return X is None or X.shape == torch.Size([])


def _call(name, X, taps, extra, out_hw, along_w):
    dt = _ffi.require_cuda_real(X, 'X')
    if X.dim() != 4:
        raise ValueError('expected a 4-D (N,C,H,W) input, got shape {}'.format(tuple(X.shape)))
    X = X.contiguous()
    N, C, H, W = X.shape
    y = X.new_empty((N, C) + tuple(out_hw))
    taps = [_ffi.host_taps(t) for t in taps]
    fn = getattr(_ffi.lib(), name + ('_f64' if dt == torch.float64 else ''))
    if N * C > 0:
        with torch.cuda.device(X.device):
            rc = fn(X.data_ptr(), y.data_ptr(), N * C, H, W, *([t.p(dt) for t in taps] + [taps[0].n] + list(extra) +
                                                                [int(along_w), _ffi.stream_of(X)]))
        _ffi.check(rc, name)
    return y


def _filter(X, h, mode, along_w):
    if _is_nothing(X):
        return torch.zeros(1, 1, 1, 1, device=X.device)                      # reference :71-72
    if isinstance(h, torch.Tensor) and h.dim() == 4:
        h = h[0]                                                              # (c,1,L,1): one copy per channel
    L = int(h.numel()) if isinstance(h, torch.Tensor) else int(np.asarray(h).size)
    ext = 2 * (L // 2) - L + 1                                                # even lengths: N + 1 outputs
    H, W = X.shape[2:]
    out = (H, W + ext) if along_w else (H + ext, W)
    return _call('b200w_dtcwt_filter', X, [h], [1 if mode == 'symmetric' else 0], out, along_w)


def colfilter(X, h, mode='symmetric'):
    """Filter the columns (dim 2) of X with ``h`` (a ``prep_filt`` tensor or a stored 1-D tap array), symmetric
    extension or zero padding; reference dtcwt/lowlevel.py:70-81."""
    return _filter(X, h, mode, False)


def rowfilter(X, h, mode='symmetric'):
    """Filter the rows (dim 3) of X with ``h``; reference dtcwt/lowlevel.py:84-94."""
    return _filter(X, h, mode, True)


def _pair(ha, hb):
    out = []
    for h in (ha, hb):
        if isinstance(h, torch.Tensor) and h.dim() == 4:
            h = h[0]
        out.append(_ffi.host_taps(h))
    if out[0].n != out[1].n:       # (the reference's conv2d would fail on the concatenated filter bank as well)
        raise ValueError('ha and hb must have the same length, got {} and {}'.format(out[0].n, out[1].n))
    if out[0].n % 2:
        raise ValueError('q-shift filters must have an even length, got {}'.format(out[0].n))
    return out


def coldfilt(X, ha, hb, highpass=False, mode='symmetric'):
    """Decimating q-shift column filter pair; reference dtcwt/lowlevel.py:97-122."""
    if _is_nothing(X):
        return torch.zeros(1, 1, 1, 1, device=X.device)
    r, c = X.shape[2:]
    if r % 4 != 0:
        raise ValueError('No. of rows in X must be a multiple of 4\n' + 'X was {}'.format(X.shape))
    if mode != 'symmetric':
        raise NotImplementedError()
    return _call('b200w_dtcwt_dfilt', X, _pair(ha, hb), [int(bool(highpass))], (r // 2, c), False)


def rowdfilt(X, ha, hb, highpass=False, mode='symmetric'):
    """Decimating q-shift row filter pair; reference dtcwt/lowlevel.py:125-151."""
    if _is_nothing(X):
        return torch.zeros(1, 1, 1, 1, device=X.device)
    r, c = X.shape[2:]
    if c % 4 != 0:
        raise ValueError('No. of cols in X must be a multiple of 4\n' + 'X was {}'.format(X.shape))
    if mode != 'symmetric':
        raise NotImplementedError()
    return _call('b200w_dtcwt_dfilt', X, _pair(ha, hb), [int(bool(highpass))], (r, c // 2), True)


def colifilt(X, ha, hb, highpass=False, mode='symmetric'):
    """Interpolating q-shift column filter pair; reference dtcwt/lowlevel.py:154-196."""
    if _is_nothing(X):
        return torch.zeros(1, 1, 1, 1, device=X.device)
    r, c = X.shape[2:]
    if r % 2 != 0:
        raise ValueError('No. of rows in X must be a multiple of 2.\n' + 'X was {}'.format(X.shape))
    if mode != 'symmetric':
        raise NotImplementedError()
    return _call('b200w_dtcwt_ifilt', X, _pair(ha, hb), [int(bool(highpass))], (2 * r, c), False)


def rowifilt(X, ha, hb, highpass=False, mode='symmetric'):
    """Interpolating q-shift row filter pair; reference dtcwt/lowlevel.py:199-239."""
    if _is_nothing(X):
        return torch.zeros(1, 1, 1, 1, device=X.device)
    r, c = X.shape[2:]
    if c % 2 != 0:
        raise ValueError('No. of cols in X must be a multiple of 2.\n' + 'X was {}'.format(X.shape))
    if mode != 'symmetric':
        raise NotImplementedError()
    return _call('b200w_dtcwt_ifilt', X, _pair(ha, hb), [int(bool(highpass))], (r, 2 * c), True)


def q2c(y, dim=-1):
    """Quads -> the two complex subimages ((re, im), (re, im)); reference dtcwt/lowlevel.py:243-259."""
    y = y / np.sqrt(2)
    a, b = y[:, :, 0::2, 0::2], y[:, :, 0::2, 1::2]
    c, d = y[:, :, 1::2, 0::2], y[:, :, 1::2, 1::2]
    return ((a - d, b + c), (a + d, b - c))


def c2q(w1, w2):
    """Two complex subimages -> real quads; reference dtcwt/lowlevel.py:262-295."""
    w1r, w1i = w1
    w2r, w2i = w2
    b, ch, r, c = w1r.shape
    y = w1r.new_zeros((b, ch, r * 2, c * 2))
    y[:, :, ::2, ::2] = w1r + w2r
    y[:, :, ::2, 1::2] = w1i + w2i
    y[:, :, 1::2, ::2] = w1i - w2i
    y[:, :, 1::2, 1::2] = -w1r + w2r
    y /= np.sqrt(2)
    return y
