"""Pin the CPU oracle (oracle/wave_oracle.c) against golden vectors produced by the reference
itself (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util


@pytest.mark.parametrize('name', util.fixtures('dwt_'))
def test_dwt_forward_inverse(name):
    g = util.load(name)
    J, mode = int(g['J']), str(g['mode'])
    filts = [g[k] for k in ('h0_col', 'h1_col', 'h0_row', 'h1_row')]
    yl, yh = orc.dwt_forward(g['x'], filts, J, mode)
    util.assert_close(yl, g['yl'], 2e-6, 'yl')
    fracs = [util.bit_equal_fraction(yl, g['yl'])]
    for j in range(J):
        util.assert_close(yh[j], g['yh%d' % j], 2e-6, 'yh%d' % j)
        fracs.append(util.bit_equal_fraction(yh[j], g['yh%d' % j]))
    if mode != 'periodization':
        # stored-tap-order FMA accumulation reproduces the reference CPU result bit for bit
        # (the periodization fold adds two partial sums in the reference, so only close there)
        assert min(fracs) > 0.97, fracs
    gf = [g[k] for k in ('g0_col', 'g1_col', 'g0_row', 'g1_row')]
    y = orc.dwt_inverse(g['yl'], [g['yh%d' % j] for j in range(J)], gf, mode)
    util.assert_close(y, g['y'], 3e-6, 'inverse')
    yh_drop = [g['yh%d' % j] for j in range(J)]
    if J > 1:
        yh_drop[0] = None
    util.assert_close(orc.dwt_inverse(g['yl'], yh_drop, gf, mode), g['y_drop0'], 3e-6, 'inverse (None highs)')


def test_config1_bit_exact():
    """BASELINE.json configs[0]: DWTForward J=1 db4 zero on randn(4,3,64,64) -- bit check."""
    g = util.load('dwt_c1_db4_zero_J1')
    filts = [g[k] for k in ('h0_col', 'h1_col', 'h0_row', 'h1_row')]
    yl, yh = orc.dwt_forward(g['x'], filts, 1, 'zero')
    assert np.array_equal(yl, g['yl'])
    assert np.array_equal(yh[0], g['yh0'])


@pytest.mark.parametrize('name', util.fixtures('dtcwt_'))
def test_dtcwt_forward_inverse(name):
    g = util.load(name)
    J, mode = int(g['J']), str(g['mode'])
    o_dim, ri_dim = int(g['o_dim']), int(g['ri_dim'])
    skip = [bool(s) for s in g['skip']]
    yl, yh = orc.dtcwt_forward(g['x'], (g['h0o'], g['h1o']), (g['h0a'], g['h0b'], g['h1a'], g['h1b']), J,
                               skip_hps=skip, o_dim=o_dim, ri_dim=ri_dim, mode=mode)
    # same tap order as the reference: bit-equal almost everywhere (oneDNN picks a different
    # summation order for a few shapes), always within 1e-6 of max|ref|
    util.assert_close(yl, g['yl'], 1e-6, 'yl')
    exact = name.startswith('dtcwt_a_a')  # short filters: oneDNN's order == stored-tap order
    assert not exact or util.bit_equal_fraction(yl, g['yl']) > 0.9
    for j in range(J):
        if skip[j]:
            assert yh[j] is None
        else:
            util.assert_close(yh[j], g['yh%d' % j], 1e-6, 'yh%d' % j)
            assert not exact or util.bit_equal_fraction(yh[j], g['yh%d' % j]) > 0.9
    yh_in = [None if skip[j] else g['yh%d' % j] for j in range(J)]
    y = orc.dtcwt_inverse(g['yl'], yh_in, (g['g0o'], g['g1o']), (g['g0a'], g['g0b'], g['g1a'], g['g1b']),
                          o_dim, ri_dim, mode)
    util.assert_close(y, g['y'], 1e-6, 'inverse')


@pytest.mark.parametrize('name', util.fixtures('scat_'))
def test_scat(name):
    g = util.load(name)
    mode, b = str(g['mode']), float(g['magbias'])
    z = orc.scat_layer(g['x'], (g['h0o'], g['h1o']), mode, b)
    util.assert_close(z, g['z'], 1e-6, 'z')
    z2 = orc.scat_layer(z, (g['h0o'], g['h1o']), mode, b)
    util.assert_close(z2, g['z2'], 2e-6, 'z2')


def test_fp64_matches_fp32_closely():
    g = util.load('dwt_db4_symmetric_J3_64')
    filts = [g[k] for k in ('h0_col', 'h1_col', 'h0_row', 'h1_row')]
    yl32, _ = orc.dwt_forward(g['x'], filts, 3, 'symmetric')
    yl64, _ = orc.dwt_forward(g['x'].astype(np.float64), filts, 3, 'symmetric')
    assert yl64.dtype == np.float64
    util.assert_close(yl32, yl64, 2e-6)


def test_bad_mode_raises():
    x = np.zeros((1, 1, 8, 8), np.float32)
    with pytest.raises(ValueError):
        orc.dwt_afb2d(x, [1, 1], [1, -1], [1, 1], [1, -1], 'constant')
    with pytest.raises(ValueError):
        orc.dwt_afb2d(x, [1, 1], [1, -1], [1, 1], [1, -1], 'bogus')


def test_line_and_row_forms_identical():
    """The oracle's readable per-line operators and its row-vectorised plane passes are the same
    arithmetic: outputs are bit-identical."""
    from pytorch_wavelets_b200.dtcwt._tables import TABLES
    rng = np.random.default_rng(0)
    lib = orc.lib()
    x = rng.standard_normal((2, 2, 44, 60)).astype(np.float32)
    g = util.load('dwt_db4_symmetric_J3_64')
    hf = [g[k] for k in ('h0_col', 'h1_col', 'h0_row', 'h1_row')]
    gf = [g[k] for k in ('g0_col', 'g1_col', 'g0_row', 'g1_row')]
    rev = lambda t, k: np.array(TABLES[t][k])[::-1].copy()  # noqa: E731
    l1 = (rev('near_sym_b', 'h0o'), rev('near_sym_b', 'h1o'))
    qs = tuple(rev('qshift_c', k) for k in ('h0a', 'h0b', 'h1a', 'h1b'))
    gl1 = (rev('near_sym_b', 'g0o'), rev('near_sym_b', 'g1o'))
    gqs = tuple(rev('qshift_c', k) for k in ('g0a', 'g0b', 'g1a', 'g1b'))

    def run():
        out = []
        for mode in ('zero', 'symmetric', 'reflect', 'periodic', 'periodization'):
            yl, yh = orc.dwt_forward(x, hf, 2, mode)
            out += [yl] + yh + [orc.dwt_inverse(yl, yh, gf, mode)]
        yl, yh = orc.dtcwt_forward(x, l1, qs, 3)
        out += [yl] + yh + [orc.dtcwt_inverse(yl, yh, gl1, gqs)]
        out.append(orc.scat_layer(x, l1))
        return out

    fast = run()
    lib.orc_set_line_forms(1)
    try:
        slow = run()
    finally:
        lib.orc_set_line_forms(0)
    for a, b in zip(fast, slow):
        assert np.array_equal(a, b)
