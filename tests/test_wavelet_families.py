"""Filter families beyond Daubechies (pytorch_wavelets_b200/wavelets.py): the reference takes them from PyWavelets
(`pywt.Wavelet(wave)`, reference dwt/transform2d.py:23-25; its tests use 'bior2.4', tests/test_dwt.py:37).  PyWavelets is
not installed here, so the constructed banks are pinned to its published taps where those are on record, and to the
properties every PyWavelets bank has (perfect reconstruction through the oracle's analysis / synthesis)."""
import numpy as np
import pytest

from oracle import oracle
from pytorch_wavelets_b200.wavelets import Wavelet

PUBLISHED = {   # PyWavelets dec_lo / rec_lo tables
    'sym4': ('dec_lo', [-0.07576571478927333, -0.02963552764599851, 0.49761866763201545, 0.8037387518059161,
                        0.29785779560527736, -0.09921954357684722, -0.012603967262037833, 0.0322231006040427]),
    'sym5': ('dec_lo', [0.027333068345077982, 0.029519490925774643, -0.039134249302383094, 0.1993975339773936,
                        0.7234076904024206, 0.6339789634582119, 0.01660210576452232, -0.17532808990845047,
                        -0.021101834024758855, 0.019538882735286728]),
    'bior2.2': ('dec_lo', [0.0, -0.1767766952966369, 0.3535533905932738, 1.0606601717798214, 0.3535533905932738,
                           -0.1767766952966369]),
    'bior1.3': ('dec_lo', [-0.08838834764831845, 0.08838834764831845, 0.7071067811865476, 0.7071067811865476,
                           0.08838834764831845, -0.08838834764831845]),
    'bior2.4': ('dec_lo', [0.0, 0.03314563036811942, -0.06629126073623884, -0.1767766952966369, 0.4198446513295126,
                           0.9943689110435825, 0.4198446513295126, -0.1767766952966369, -0.06629126073623884,
                           0.03314563036811942]),
    'bior3.1': ('dec_lo', [-0.3535533905932738, 1.0606601717798214, 1.0606601717798214, -0.3535533905932738]),
    'bior4.4': ('rec_lo', [0.0, -0.06453888262869706, -0.04068941760916406, 0.41809227322161724, 0.7884856164055829,
                           0.41809227322161724, -0.04068941760916406, -0.06453888262869706, 0.0, 0.0]),
    'rbio2.2': ('dec_lo', [0.0, 0.0, 0.3535533905932738, 0.7071067811865476, 0.3535533905932738, 0.0]),
}

NAMES = ['sym2', 'sym3', 'sym4', 'sym5', 'sym6', 'sym8', 'coif1', 'bior1.1', 'bior1.3', 'bior1.5', 'bior2.2', 'bior2.4',
         'bior2.6', 'bior2.8', 'bior3.1', 'bior3.3', 'bior3.5', 'bior3.7', 'bior3.9', 'bior4.4', 'rbio1.3', 'rbio2.2',
         'rbio2.4', 'rbio3.3', 'rbio4.4']


@pytest.mark.parametrize('name', sorted(PUBLISHED))
def test_published_taps(name):
    field, ref = PUBLISHED[name]
    got = np.asarray(getattr(Wavelet(name), field))
    assert got.shape == (len(ref),)
    assert np.abs(got - np.asarray(ref)).max() < 5e-12


def test_bior22_high_pass_convention():
    w = Wavelet('bior2.2')   # PyWavelets: dec_hi = [0, r, -2r, r, 0, 0], rec_hi[k] = (-1)^k dec_lo[k]
    r = 0.3535533905932738
    assert np.allclose(w.dec_hi, [0.0, r, -2 * r, r, 0.0, 0.0], atol=1e-14)
    assert np.allclose(w.rec_hi, [0.0, 0.1767766952966369, r, -1.0606601717798214, r, 0.1767766952966369], atol=1e-14)


@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('mode', ['periodization', 'zero', 'symmetric'])
def test_perfect_reconstruction_through_the_oracle(name, mode):
    w = Wavelet(name)
    L = len(w.dec_lo)
    assert L % 2 == 0 and len(w.dec_hi) == len(w.rec_lo) == len(w.rec_hi) == L
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 2, 64, 48))
    # the module stores analysis taps reversed (reference prep_filt_afb2d, dwt/lowlevel.py:925-947) and synthesis as is
    filts_a = (np.asarray(w.dec_lo)[::-1].copy(), np.asarray(w.dec_hi)[::-1].copy())
    filts_s = (np.asarray(w.rec_lo), np.asarray(w.rec_hi))
    yl, yh = oracle.dwt_forward(x, filts_a + filts_a, 2, mode)
    y = oracle.dwt_inverse(yl, yh, filts_s + filts_s, mode)
    assert y.shape == x.shape
    assert np.abs(y - x).max() < 1e-9


def test_unknown_family_still_raises_without_pywavelets():
    with pytest.raises(ValueError):
        Wavelet('coif5')
    with pytest.raises(ValueError):
        Wavelet('sym7')
