"""CPU: host-side mirror of the reference interface -- buffer names/shapes/values, aliases, errors."""
import numpy as np
import pytest
import torch

import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import wavelets
from pytorch_wavelets_b200.dtcwt import coeffs
from pytorch_wavelets_b200.dtcwt.transform_funcs import get_dimensions5, get_dimensions6, highs_shape_strides
from pytorch_wavelets_b200.dwt import lowlevel
from tests import util

DB4 = [0.23037781330885523, 0.7148465705525415, 0.6308807679295904, -0.02798376941698385,
       -0.18703481171888114, 0.030841381835986965, 0.032883011666982945, -0.010597401784997278]


def test_db4_matches_published_taps():
    w = wavelets.Wavelet('db4')
    assert np.abs(np.array(w.rec_lo) - DB4).max() < 1e-12
    assert np.allclose(w.dec_lo, DB4[::-1], atol=1e-12)
    assert np.allclose(w.rec_hi, [(-1) ** k * w.dec_lo[k] for k in range(8)])
    assert np.allclose(w.dec_hi, w.rec_hi[::-1])


@pytest.mark.parametrize('N', range(1, 21))
def test_daubechies_orthonormal_with_vanishing_moments(N):
    h = wavelets.daubechies(N)
    assert len(h) == 2 * N
    assert abs(h.sum() - np.sqrt(2)) < 1e-9
    g = np.array([(-1) ** k * h[::-1][k] for k in range(2 * N)])
    k = np.arange(2 * N)
    for p in range(min(N, 5)):  # higher moments cancel catastrophically in float64 for long filters
        assert abs((g * k ** p).sum()) < 1e-6 * max(1.0, float((np.abs(g) * k ** p).sum()))


def test_dwt_coeff_len():
    assert wavelets.dwt_coeff_len(512, 8, 'symmetric') == 259
    assert wavelets.dwt_coeff_len(127, 8, 'periodization') == 64
    assert wavelets.dwt_coeff_len(127, 8, 'per') == 64


@pytest.mark.parametrize('name', util.fixtures('dwt_'))
def test_dwt_buffers_match_reference(name):
    g = util.load(name)
    f = pw.DWTForward(J=int(g['J']), wave=str(g['wave']), mode=str(g['mode']))
    i = pw.DWTInverse(wave=str(g['wave']), mode=str(g['mode']))
    sd = f.state_dict()
    assert list(sd.keys()) == ['h0_col', 'h1_col', 'h0_row', 'h1_row']
    L = g['h0_col'].size
    assert sd['h0_col'].shape == (1, 1, L, 1) and sd['h0_row'].shape == (1, 1, 1, L)
    for k in sd:
        assert np.abs(sd[k].numpy().ravel() - g[k]).max() < 1e-6
    sd = i.state_dict()
    assert list(sd.keys()) == ['g0_col', 'g1_col', 'g0_row', 'g1_row']
    for k in sd:
        assert np.abs(sd[k].numpy().ravel() - g[k]).max() < 1e-6


@pytest.mark.parametrize('name', util.fixtures('dtcwt_'))
def test_dtcwt_buffers_match_reference(name):
    g = util.load(name)
    f = pw.DTCWTForward(biort=str(g['biort']), qshift=str(g['qshift']), J=int(g['J']))
    i = pw.DTCWTInverse(biort=str(g['biort']), qshift=str(g['qshift']))
    sd = f.state_dict()
    assert list(sd.keys()) == ['h0o', 'h1o', 'h0a', 'h0b', 'h1a', 'h1b']
    for k in sd:
        assert sd[k].shape == (1, 1, g[k].size, 1)
        assert np.array_equal(sd[k].numpy().ravel(), g[k])
    sd = i.state_dict()
    assert list(sd.keys()) == ['g0o', 'g1o', 'g0a', 'g0b', 'g1a', 'g1b']
    for k in sd:
        assert np.array_equal(sd[k].numpy().ravel(), g[k])


def test_scat_parameters():
    s = pw.ScatLayer()
    names = [n for n, _ in s.named_parameters()]
    assert names == ['h0o', 'h1o']
    assert all(not p.requires_grad for p in s.parameters())
    assert 'near_sym_a' in repr(s)


def test_aliases_and_exports():
    assert pw.DWT is pw.DWTForward and pw.IDWT is pw.DWTInverse
    assert pw.DWT2D is pw.DWTForward and pw.IDWT2D is pw.DWTInverse
    assert pw.DTCWT is pw.DTCWTForward and pw.IDTCWT is pw.DTCWTInverse
    for n in pw.__all__:
        assert hasattr(pw, n)


def test_constructor_errors():
    with pytest.raises(ValueError):
        pw.DTCWTForward(o_dim=2, ri_dim=2)
    with pytest.raises(ValueError):
        lowlevel.mode_to_int('bogus')
    with pytest.raises(ValueError):
        lowlevel.int_to_mode(9)
    with pytest.raises(ValueError):
        pw.DWTForward(wave='sym-not-installed-4')
    assert [lowlevel.mode_to_int(m) for m in
            ('zero', 'symmetric', 'periodization', 'constant', 'reflect', 'replicate', 'periodic')] == list(range(7))
    assert lowlevel.mode_to_int('per') == 2


def test_filter_tuples_follow_reference_quirk():
    a0, a1 = np.arange(4.) + 1, np.arange(4.) + 5
    b0, b1 = np.arange(6.) + 10, np.arange(6.) + 20
    f = pw.DWTForward(wave=(a0, a1, b0, b1))
    assert f.h0_col.shape == (1, 1, 4, 1) and f.h0_row.shape == (1, 1, 1, 6)
    assert np.array_equal(f.h0_col.numpy().ravel(), a0[::-1])  # stored reversed
    f2 = pw.DWTForward(wave=(a0, a1))
    assert np.array_equal(f2.h1_row.numpy().ravel(), a1[::-1])


def test_cpu_tensor_is_rejected_not_silently_computed():
    with pytest.raises(NotImplementedError):
        pw.DWTForward()(torch.zeros(1, 1, 8, 8))
    with pytest.raises(NotImplementedError):
        pw.DTCWTForward()(torch.zeros(1, 1, 8, 8))
    with pytest.raises(NotImplementedError):
        pw.ScatLayer()(torch.zeros(1, 1, 8, 8))


def test_coeff_tables():
    h0o, g0o, h1o, g1o = coeffs.biort('near_sym_a')
    assert h0o.shape == (5, 1) and h1o.shape == (7, 1) and g0o.shape == (7, 1) and g1o.shape == (5, 1)
    q = coeffs.qshift('qshift_a')
    assert len(q) == 8 and all(f.shape == (10, 1) for f in q)
    assert np.allclose(q[0][::-1], q[1])  # b tree = time reverse of a tree
    with pytest.raises(IOError):
        coeffs.biort('nope')
    with pytest.raises(ValueError):
        coeffs.qshift('near_sym_a')


@pytest.mark.parametrize('o_dim,ri_dim', [(2, -1), (1, 2), (4, 5), (3, 1), (5, 2), (2, 3), (1, -1)])
def test_layout_helpers_agree(o_dim, ri_dim):
    o5, ri, h5, w5 = get_dimensions5(o_dim, ri_dim)
    _, _, h6, w6 = get_dimensions6(o_dim, ri_dim)
    shape, hs = highs_shape_strides(2, 3, 5, 7, o5, ri)
    assert shape[o_dim % 6] == 6 and shape[ri_dim % 6] == 2
    assert shape[h6] == 5 and shape[w6] == 7
    t = torch.empty(shape)
    names = ['n', 'c', 'o', 'h', 'w', 'r']
    want = {'n': 2, 'c': 3, 'o': 6, 'h': 5, 'w': 7, 'r': 2}
    strides = t.stride()
    for k, s in zip(names, hs):
        dims = [d for d in range(6) if strides[d] == s and shape[d] == want[k]]
        assert dims, (k, s, shape, strides)
