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

* Families beyond Daubechies, also constructed (no tables to copy, no PyWavelets offline):
  ``sym2``..``sym6``, ``sym8`` (least-asymmetric root selection of the same half-band polynomial; sym4 / sym5 agree
  with PyWavelets' published taps to 1e-12, the orientation of each is fixed to PyWavelets'), the biorthogonal spline
  family ``biorNr.Nd`` / ``rbioNr.Nd`` for (Nr.Nd) in 1.1 1.3 1.5 2.2 2.4 2.6 2.8 3.1 3.3 3.5 3.7 3.9 (Cohen-Daubechies-
  Feauveau: rec_lo = sqrt(2) ((1+z)/2)^Nr, dec_lo = sqrt(2) ((1+z)/2)^Nd P_K(y), K = (Nr+Nd)/2, zero-padded and aligned
  the way PyWavelets stores them; bior2.2, 1.3, 3.1, 3.3, 2.4 checked against its published taps) and ``bior4.4`` /
  ``rbio4.4`` (the CDF 9/7 pair: the real root of P_4 goes to rec_lo, the complex pair to dec_lo).

``coif1`` is the published 6-tap table.

If PyWavelets *is* importable it is used for every remaining family (coif2.., sym7, sym9.., bior5.5, bior6.8, dmey);
otherwise those names raise ``ValueError`` and the caller can pass explicit filter tuples, exactly as the reference allows.
"""
import math

import numpy as np

__all__ = ['Wavelet', 'dwt_coeff_len', 'daubechies', 'symlet', 'biorthogonal']

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


def _half_band_root_groups(N):
    """Roots of the Daubechies half-band polynomial P_N(y), y = (2 - z - 1/z)/4, as z-root groups: each entry is
    (roots inside the unit circle, their reciprocals outside) for one real y-root or one complex-conjugate pair."""
    from math import comb
    ys = np.roots(np.array([comb(N - 1 + k, k) for k in range(N)][::-1], dtype=np.float64))

    def zpair(y):
        b = 2.0 - 4.0 * y
        d = np.sqrt(b * b - 4.0 + 0j)
        z1, z2 = (b + d) / 2.0, (b - d) / 2.0
        return (z1, z2) if abs(z1) < 1.0 else (z2, z1)

    used = [False] * len(ys)
    groups = []
    for i, y in enumerate(ys):
        if used[i]:
            continue
        used[i] = True
        if abs(y.imag) < 1e-9:
            zi, zo = zpair(y.real + 0j)
            groups.append(([zi], [zo]))
        else:
            j = min((k for k in range(len(ys)) if not used[k]), key=lambda k: abs(ys[k] - np.conj(y)))
            used[j] = True
            a, b = zpair(y), zpair(ys[j])
            groups.append(([a[0], b[0]], [a[1], b[1]]))
    return groups


def _phase_nonlinearity(h):
    w = np.linspace(0.0, np.pi, 513)[1:-1]
    k = np.arange(len(h))
    H = np.exp(-1j * np.outer(w, k)) @ h
    ph = np.unwrap(np.angle(H))
    A = np.vstack([w, np.ones_like(w)]).T
    wt = np.abs(H)
    coef = np.linalg.lstsq(A * wt[:, None], ph * wt, rcond=None)[0]
    return float(np.sum((wt * (ph - A @ coef)) ** 2))


# symN whose construction below is pinned to PyWavelets' taps (value = True: PyWavelets stores the time reverse of the
# polynomial-order filter).  sym7 and sym9.. use a root selection this criterion does not reproduce: not offered.
_SYM_REVERSED = {2: False, 3: False, 4: False, 5: True, 6: False, 8: False}


def symlet(N):
    """``dec_lo`` (length 2N) of the least-asymmetric Daubechies wavelet symN: the inside/outside choice per root group
    of the half-band polynomial that minimises the (magnitude-weighted) deviation of the phase from linear."""
    if N not in _SYM_REVERSED:
        raise ValueError('sym%d is not built in (sym2..sym6, sym8 are)' % N)
    key = ('sym', N)
    if key in _CACHE:
        return _CACHE[key].copy()
    if N <= 3:
        h = daubechies(N)[::-1].copy()            # sym2 = db2, sym3 = db3
    else:
        import itertools
        groups = _half_band_root_groups(N)
        best = None
        for sel in itertools.product((0, 1), repeat=len(groups)):
            zr = [z for g, c in zip(groups, sel) for z in g[c]]
            q = np.real(np.poly(np.array(zr)))
            for _ in range(N):
                q = np.convolve(q, [1.0, 1.0])
            q = q * (math.sqrt(2.0) / q.sum())
            m = _phase_nonlinearity(q)
            if best is None or m < best[0] - 1e-12:
                best = (m, q)
        h = best[1][::-1].copy() if _SYM_REVERSED[N] else best[1]
    L = 2 * N
    for s_ in range(N):
        v = float(np.dot(h[:L - 2 * s_], h[2 * s_:]))
        if abs(v - (1.0 if s_ == 0 else 0.0)) > 1e-8:
            raise RuntimeError('sym%d construction failed orthonormality' % N)
    _CACHE[key] = h
    return h.copy()


_BIOR_SPLINE = {(1, 1), (1, 3), (1, 5), (2, 2), (2, 4), (2, 6), (2, 8), (3, 1), (3, 3), (3, 5), (3, 7), (3, 9)}


def _spline(n):
    h = np.array([1.0])
    for _ in range(n):
        h = np.convolve(h, [0.5, 0.5])
    return h


def _P_of_z(K):
    """P_K(y) = sum_k C(K-1+k, k) y^k as a symmetric polynomial in z (2K-1 coefficients), y = (2 - z - 1/z)/4."""
    from math import comb
    out = np.zeros(2 * K - 1)
    yk = np.array([1.0])
    for k in range(K):
        pad = (2 * K - 1 - len(yk)) // 2
        out[pad:pad + len(yk)] += comb(K - 1 + k, k) * yk
        yk = np.convolve(yk, [-0.25, 0.5, -0.25])
    return out


def biorthogonal(Nr, Nd):
    """(dec_lo, dec_hi, rec_lo, rec_hi) of PyWavelets' ``bior<Nr>.<Nd>`` as float64 arrays of equal (even) length."""
    key = ('bior', Nr, Nd)
    if key in _CACHE:
        return tuple(a.copy() for a in _CACHE[key])
    if (Nr, Nd) in _BIOR_SPLINE:
        K = (Nr + Nd) // 2
        rec = math.sqrt(2.0) * _spline(Nr)
        dec = math.sqrt(2.0) * np.convolve(_spline(Nd), _P_of_z(K))
    elif (Nr, Nd) == (4, 4):
        # CDF 9/7: P_4 has one real root (-> the 7-tap synthesis low-pass) and one complex pair (-> the 9-tap analysis one)
        from math import comb
        ys = np.roots(np.array([comb(3 + k, k) for k in range(4)][::-1], dtype=np.float64))
        real = [y for y in ys if abs(y.imag) < 1e-9]
        cplx = [y for y in ys if abs(y.imag) >= 1e-9]

        def factor(yr):   # prod (y - y_r) as a polynomial in z
            f = np.array([1.0 + 0j])
            for y0 in yr:
                f = np.convolve(f, np.array([-0.25, 0.5 - y0, -0.25]))
            return np.real(f)
        rec = np.convolve(_spline(4), factor(real))
        dec = np.convolve(_spline(4), factor(cplx))
        rec = rec * (math.sqrt(2.0) / rec.sum())
        dec = dec * (math.sqrt(2.0) / dec.sum())
    else:
        raise ValueError('bior%d.%d is not built in' % (Nr, Nd))
    L = max(len(dec), len(rec))
    L += L & 1
    dl, rl = np.zeros(L), np.zeros(L)
    dl[L - len(dec):] = dec                         # PyWavelets right-aligns the analysis low-pass ...
    cd = (L - len(dec)) + (len(dec) - 1) / 2.0
    cr = cd - 1.0 if (len(dec) & 1) else cd         # ... and centres the synthesis one a sample earlier (odd lengths)
    s0 = int(round(cr - (len(rec) - 1) / 2.0))
    rl[s0:s0 + len(rec)] = rec
    k = np.arange(L)
    rh = ((-1.0) ** k) * dl
    dh = ((-1.0) ** (k + 1)) * rl
    _CACHE[key] = (dl, dh, rl, rh)
    return tuple(a.copy() for a in _CACHE[key])


# coif1 (Daubechies' 6-tap coiflet), published dec_lo; orthonormal to 1e-12 with two vanishing wavelet moments
_COIF1_DEC_LO = (-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554,
                 0.3378976624578092, -0.0727326195128539)


def _builtin_bank(lname):
    """Filter bank of a built-in (non-dbN) name, or None."""
    if lname == 'coif1':
        dec_lo = list(_COIF1_DEC_LO)
        rec_hi = [((-1.0) ** k) * dec_lo[k] for k in range(6)]
        return dec_lo, rec_hi[::-1], dec_lo[::-1], rec_hi
    if lname.startswith('sym') and lname[3:].isdigit() and int(lname[3:]) in _SYM_REVERSED:
        d = symlet(int(lname[3:]))
        dec_lo = d.tolist()
        rec_lo = d[::-1].tolist()
        rec_hi = [((-1.0) ** k) * dec_lo[k] for k in range(len(dec_lo))]
        return dec_lo, rec_hi[::-1], rec_lo, rec_hi
    for fam in ('bior', 'rbio'):
        if lname.startswith(fam):
            try:
                nr, nd = (int(t) for t in lname[4:].split('.'))
            except ValueError:
                return None
            if (nr, nd) not in _BIOR_SPLINE and (nr, nd) != (4, 4):
                return None
            dl, dh, rl, rh = biorthogonal(nr, nd)
            if fam == 'bior':
                return dl.tolist(), dh.tolist(), rl.tolist(), rh.tolist()
            # reverse biorthogonal: analysis and synthesis banks swapped and time-reversed
            return rl[::-1].tolist(), rh[::-1].tolist(), dl[::-1].tolist(), dh[::-1].tolist()
    return None


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
            elif _builtin_bank(lname) is not None:
                dec_lo, dec_hi, rec_lo, rec_hi = _builtin_bank(lname)
                N = None
            else:
                try:
                    import pywt  # noqa: optional dependency, same as the reference
                except ImportError:
                    raise ValueError(
                        "wavelet %r needs PyWavelets (built in: haar, db1..db20, sym2..sym6, sym8, coif1, bior/rbio spline pairs, "
                        "bior4.4); "
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
