"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# fp32 parity tolerance, relative to max|ref| of the tensor compared.  Same-algorithm fp32
# differences (FMA contraction, summation order of the transposed convolutions, x*(1/sqrt2) vs
# x/sqrt2) are ~1e-7..1e-6; the reference's own tests use decimal=3..4.
RTOL_F32 = 1e-5


def fixtures(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + '*.npz')))


def load(name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return {k: d[k] for k in d.files}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def bit_equal_fraction(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float((a == b).mean())


def assert_close(a, b, tol=RTOL_F32, what=''):
    e = rel_err(a, b)
    assert e <= tol, '%s: rel err %.3e > %.1e' % (what, e, tol)
    return e
