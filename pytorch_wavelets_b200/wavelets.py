"""Wavelet filter banks and coefficient-length rule for the 2-D DWT path.

The reference obtains both from PyWavelets (third-party, not vendored and not
installed in this image): ``pywt.Wavelet(name)`` at reference
``pytorch_wavelets/dwt/transform2d.py:23-25,92-94`` and
``pywt.dwt_coeff_len`` at ``pytorch_wavelets/dwt/lowlevel.py:153``.  This module
supplies the same two things without PyWavelets:

* ``Wavelet(name)`` for the orthogonal Daubechies family ``haar``/``db1``..``db20``,
  built from the published construction (spectral factorisation of the
  Daubechies half-band polynomial, minimum-phase root choice -> ``rec_lo``;
  ``dec_lo = rec_lo[::-1]``, ``rec_hi[k] = (-1)^k dec_lo[k]``,
  ``dec_hi = rec_hi[::-1]``), checked at import of each wavelet for
  orthonormality.  db4 generated this way agrees with PyWavelets' table to
  < 1e-12 (tests/test_wavelets.py).
* ``dwt_coeff_len(data_len, filter_len, mode)``.

If PyWavelets *is* importable it is used for every other family (sym, coif,
bior, ...); otherwise those names raise ``ValueError`` and the caller can pass
explicit filter tuples, exactly as the reference allows.
"""
import math

import numpy as np

__all__ = ['Wavelet', 'dwt_coeff_len', 'daubechies']

_CACHE = {}


def daubechies(N):
    """Return ``rec_lo`` (length 2N, float64) of the Daubechies wavelet dbN."""
    if N < 1 or N > 20:
        raise ValueError('dbN supported for 1 <= N <= 20, got %d' % N)
    if N in _CACHE:
        return _CACHE[N].copy()
    if N == 1:
        h = np.array([1.0, 1.0]) / math.sqrt(2.0)
        _CACHE[N] = h
        return h.copy()
    # P(y) = sum_k C(N-1+k, k) y^k, with y = (2 - z - 1/z)/4.  Work with the
    # roots in y (well conditioned), then map each y-root to the z-root pair
    # z + 1/z = 2 - 4y and keep the one inside the unit circle.
    from math import comb
    coeffs_y = [comb(N - 1 + k, k) for k in range(N)]       # ascending in y
    yroots = np.roots(np.array(coeffs_y[::-1], dtype=np.float64))
    zroots = []
    for y in yroots:
        b = 2.0 - 4.0 * y                                    # z^2 - b z + 1 = 0
        d = np.sqrt(b * b - 4.0 + 0j)
        z1 = (b + d) / 2.0
        z2 = (b - d) / 2.0
        zroots.append(z1 if abs(z1) < 1.0 else z2)
    q = np.real(np.poly(np.array(zroots)))                   # prod (z - r_i)
    h = q
    for _ in range(N):                                       # times (1 + z)^N
        h = np.convolve(h, [1.0, 1.0])
    h = h * (math.sqrt(2.0) / h.sum())
    # orthonormality check (double-shift orthogonality)
    L = 2 * N
    for s in range(N):
        v = float(np.dot(h[:L - 2 * s], h[2 * s:]))
        tgt = 1.0 if s == 0 else 0.0
        if abs(v - tgt) > 1e-8:
            raise RuntimeError('db%d construction failed orthonormality (%g)' % (N, v - tgt))
    _CACHE[N] = h
    return h.copy()


class Wavelet(object):
    """Minimal stand-in for ``pywt.Wavelet``: ``dec_lo, dec_hi, rec_lo, rec_hi``."""

    def __init__(self, name, filter_bank=None):
        self.name = name
        if filter_bank is not None:
            dec_lo, dec_hi, rec_lo, rec_hi = [list(map(float, f)) for f in filter_bank]
        else:
            lname = name.lower()
            if lname == 'haar':
                N = 1
            elif lname.startswith('db') and lname[2:].isdigit():
                N = int(lname[2:])
            else:
                try:
                    import pywt  # noqa: optional dependency, same as the reference
                except ImportError:
                    raise ValueError(
                        "wavelet %r needs PyWavelets (only haar/db1..db20 are built in); "
                        "pass explicit filter tuples instead" % (name,))
                w = pywt.Wavelet(name)
                dec_lo, dec_hi, rec_lo, rec_hi = w.dec_lo, w.dec_hi, w.rec_lo, w.rec_hi
                N = None
            if N is not None:
                r = daubechies(N)
                rec_lo = r.tolist()
                dec_lo = r[::-1].tolist()
                rec_hi = [((-1.0) ** k) * dec_lo[k] for k in range(len(dec_lo))]
                dec_hi = rec_hi[::-1]
        self.dec_lo, self.dec_hi = list(dec_lo), list(dec_hi)
        self.rec_lo, self.rec_hi = list(rec_lo), list(rec_hi)
        self.dec_len = len(self.dec_lo)
        self.rec_len = len(self.rec_lo)

    @property
    def filter_bank(self):
        return (self.dec_lo, self.dec_hi, self.rec_lo, self.rec_hi)

    def __repr__(self):
        return 'Wavelet(%r)' % (self.name,)


def dwt_coeff_len(data_len, filter_len, mode):
    """``pywt.dwt_coeff_len``: ceil(N/2) for periodization else floor((N+L-1)/2)."""
    if data_len < 1:
        raise ValueError('Value of data_len must be greater than zero.')
    if filter_len < 1:
        raise ValueError('Value of filter_len must be greater than zero.')
    if mode in ('per', 'periodization'):
        return (data_len + 1) // 2
    return (data_len + filter_len - 1) // 2
