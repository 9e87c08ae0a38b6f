"""DTCWT filter tables by name -- same functions and return order as the reference
``pytorch_wavelets/dtcwt/coeffs.py`` (``biort`` :34-38, ``level1`` :41-77, ``qshift`` :80-116), reading the
taps from the generated ``_tables`` module instead of npz resources.  Filters are float64 column vectors."""
import numpy as np

from pytorch_wavelets_b200.dtcwt._tables import TABLES


def _load(name, varnames):
    try:
        tab = TABLES[name]
    except KeyError:
        raise IOError('No such wavelet: {0}'.format(name))
    try:
        return tuple(np.array(tab[k], dtype=np.float64).reshape(-1, 1) for k in varnames)
    except KeyError:
        raise ValueError('Wavelet does not define ({0}) coefficients'.format(', '.join(varnames)))


def level1(name, compact=False):
    """h0o, g0o, h1o, g1o (compact=True; plus h2o, g2o for near_sym_b_bp), else the 8 tree filters."""
    if compact:
        if name == 'near_sym_b_bp':
            return _load(name, ('h0o', 'g0o', 'h1o', 'g1o', 'h2o', 'g2o'))
        return _load(name, ('h0o', 'g0o', 'h1o', 'g1o'))
    return _load(name, ('h0a', 'h0b', 'g0a', 'g0b', 'h1a', 'h1b', 'g1a', 'g1b'))


def biort(name):
    return level1(name, compact=True)


def qshift(name):
    """h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b (plus the four *2* filters for qshift_b_bp)."""
    if name == 'qshift_b_bp':
        return _load(name, ('h0a', 'h0b', 'g0a', 'g0b', 'h1a', 'h1b', 'g1a', 'g1b', 'h2a', 'h2b', 'g2a', 'g2b'))
    return _load(name, ('h0a', 'h0b', 'g0a', 'g0b', 'h1a', 'h1b', 'g1a', 'g1b'))
