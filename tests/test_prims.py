"""The reference's standalone 1-D DTCWT primitives (dtcwt/lowlevel.py:70-239: colfilter / rowfilter -- including EVEN
filter lengths, which give N + 1 outputs, reference tests/test_colfilter.py:52-61 -- coldfilt / rowdfilt, colifilt /
rowifilt).  Golden vectors: outputs of the unmodified reference (tests/golden/make_golden.py prims).  CPU: the oracle's
restatement against them; GPU: pytorch_wavelets_b200.dtcwt.lowlevel (csrc/k_prims.cu) against them and bit-for-bit
against the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import util

G = util.load('prims_24x28')
FILT = [(L, ax, mode) for L in (5, 7, 4, 6, 2) for ax in ('col', 'row') for mode in ('symmetric', 'zero')]
QS = [(q, op, hp) for q in ('qshift_a', 'qshift_b', 'qshift_c') for op in ('coldfilt', 'rowdfilt', 'colifilt', 'rowifilt')
      for hp in (0, 1)]


@pytest.mark.parametrize('L,ax,mode', FILT)
def test_oracle_filter_vs_reference(L, ax, mode):
    y = orc.filter1d(G['x'], G['filt%d_h' % L], symmetric=(mode == 'symmetric'), along_w=(ax == 'row'))
    ref = G['filt%d_%s_%s' % (L, ax, mode)]
    assert y.shape == ref.shape          # even lengths: one more output along the filtered dimension
    util.assert_close(y, ref, 1e-6, 'filter')


@pytest.mark.parametrize('q,op,hp', QS)
def test_oracle_qshift_primitives_vs_reference(q, op, hp):
    fn = orc.dfilt1d if 'dfilt' in op else orc.ifilt1d
    y = fn(G['x'], G[q + '_ha'], G[q + '_hb'], highpass=bool(hp), along_w=op.startswith('row'))
    util.assert_close(y, G['%s_%s_%d' % (q, op, hp)], 1e-6, op)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to('cuda')


@pytest.mark.gpu
@pytest.mark.parametrize('L,ax,mode', FILT)
def test_gpu_filter_vs_reference_and_oracle(L, ax, mode):
    from pytorch_wavelets_b200.dtcwt import lowlevel as ll
    h = ll.prep_filt(G['filt%d_h' % L][::-1].copy(), 3)     # prep_filt reverses: hand it the un-reversed filter
    assert np.array_equal(h[0].numpy().ravel(), G['filt%d_h' % L].astype(np.float32))
    fn = ll.colfilter if ax == 'col' else ll.rowfilter
    y = fn(_t(G['x']), h.to('cuda'), mode=mode).cpu().numpy()
    ref = G['filt%d_%s_%s' % (L, ax, mode)]
    assert y.shape == ref.shape
    util.assert_close(y, ref, 1e-6, 'filter vs reference')
    o = orc.filter1d(G['x'], G['filt%d_h' % L], symmetric=(mode == 'symmetric'), along_w=(ax == 'row'))
    assert np.array_equal(y, o)


@pytest.mark.gpu
@pytest.mark.parametrize('q,op,hp', QS)
def test_gpu_qshift_primitives_vs_reference_and_oracle(q, op, hp):
    from pytorch_wavelets_b200.dtcwt import lowlevel as ll
    ha, hb = _t(G[q + '_ha']), _t(G[q + '_hb'])
    y = getattr(ll, op)(_t(G['x']), ha, hb, highpass=bool(hp)).cpu().numpy()
    util.assert_close(y, G['%s_%s_%d' % (q, op, hp)], 1e-6, op)
    fn = orc.dfilt1d if 'dfilt' in op else orc.ifilt1d
    assert np.array_equal(y, fn(G['x'], G[q + '_ha'], G[q + '_hb'], highpass=bool(hp), along_w=op.startswith('row')))


@pytest.mark.gpu
def test_gpu_primitives_f64_errors_and_q2c_roundtrip():
    from pytorch_wavelets_b200.dtcwt import lowlevel as ll
    x = torch.randn(1, 2, 16, 20, dtype=torch.float64, device='cuda')
    h = torch.tensor(G['filt4_h'], dtype=torch.float64)
    y = ll.colfilter(x, h)
    assert y.dtype == torch.float64 and tuple(y.shape) == (1, 2, 17, 20)
    o = orc.filter1d(x.cpu().numpy(), G['filt4_h'].astype(np.float64), symmetric=True, along_w=False)
    util.assert_close(y.cpu().numpy(), o, 1e-13, 'f64 filter')
    with pytest.raises(ValueError):
        ll.coldfilt(torch.randn(1, 1, 18, 16, device='cuda'), _t(G['qshift_a_ha']), _t(G['qshift_a_hb']))
    with pytest.raises(ValueError):
        ll.rowifilt(torch.randn(1, 1, 16, 15, device='cuda'), _t(G['qshift_a_ha']), _t(G['qshift_a_hb']))
    with pytest.raises(NotImplementedError):
        ll.rowdfilt(torch.randn(1, 1, 16, 16, device='cuda'), _t(G['qshift_a_ha']), _t(G['qshift_a_hb']), mode='zero')
    with pytest.raises(ValueError):     # filter pair of unequal length
        ll.coldfilt(torch.randn(1, 1, 16, 16, device='cuda'), _t(G['qshift_a_ha']), _t(G['qshift_b_hb']))
    q = torch.randn(1, 2, 8, 12, device='cuda')
    w1, w2 = ll.q2c(q)
    assert float((ll.c2q(w1, w2) - q).abs().max()) < 1e-6
