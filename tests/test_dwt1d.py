"""1-D DWT.  CPU: the oracle against the reference's golden vectors (tests/golden/dwt1d_*.npz).
GPU (-m gpu): DWT1DForward / DWT1DInverse on the CUDA kernels against the goldens and the oracle (analysis bit-identical
to the oracle), None band-passes, odd lengths, the gradient identities of the reference's tests."""
import numpy as np
import pytest
import torch

import pytorch_wavelets_b200 as pw
from oracle import oracle as orc
from tests import util


@pytest.mark.parametrize('name', util.fixtures('dwt1d_'))
def test_oracle_matches_reference_1d(name):
    g = util.load(name)
    J, mode = int(g['J']), str(g['mode'])
    yl, yh = orc.dwt1d_forward(g['x'], (g['h0'], g['h1']), J, mode)
    util.assert_close(yl, g['yl'], util.RTOL_F32, 'yl')
    for j, a in enumerate(yh):
        util.assert_close(a, g['yh%d' % j], util.RTOL_F32, 'yh%d' % j)
    hs = [g['yh%d' % j] for j in range(J)]
    util.assert_close(orc.dwt1d_inverse(g['yl'], hs, (g['g0'], g['g1']), mode), g['y'], util.RTOL_F32, 'inverse')
    util.assert_close(orc.dwt1d_inverse(g['yl'], [None] + hs[1:], (g['g0'], g['g1']), mode), g['y_drop0'], util.RTOL_F32,
                      'inverse None')


def test_module_buffers_and_errors_cpu():
    f = pw.DWT1DForward(J=2, wave='db3', mode='symmetric')
    assert tuple(f.h0.shape) == (1, 1, 6) and tuple(f.h1.shape) == (1, 1, 6)
    i = pw.DWT1DInverse(wave='db3')
    assert tuple(i.g0.shape) == (1, 1, 6)
    assert pw.DWT1D is pw.DWT1DForward and pw.IDWT1D is pw.DWT1DInverse
    with pytest.raises(NotImplementedError):
        f(torch.zeros(1, 1, 32))          # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        f(torch.zeros(1, 1, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize('name', util.fixtures('dwt1d_'))
def test_dwt1d_golden_gpu(name):
    g = util.load(name)
    J, mode, wave = int(g['J']), str(g['mode']), str(g['wave'])
    f = pw.DWT1DForward(J=J, wave=wave, mode=mode).cuda()
    i = pw.DWT1DInverse(wave=wave, mode=mode).cuda()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    yl, yh = f(t(g['x']))
    util.assert_close(yl.cpu().numpy(), g['yl'], util.RTOL_F32, 'yl')
    oyl, oyh = orc.dwt1d_forward(g['x'], (g['h0'], g['h1']), J, mode)
    assert np.array_equal(yl.cpu().numpy(), oyl)                                # same FMA order as the oracle
    for j in range(J):
        util.assert_close(yh[j].cpu().numpy(), g['yh%d' % j], util.RTOL_F32, 'yh%d' % j)
        assert np.array_equal(yh[j].cpu().numpy(), oyh[j])
    hs = [t(g['yh%d' % j]) for j in range(J)]
    util.assert_close(i((t(g['yl']), hs)).cpu().numpy(), g['y'], util.RTOL_F32, 'inverse')
    util.assert_close(i((t(g['yl']), [None] + hs[1:])).cpu().numpy(), g['y_drop0'], util.RTOL_F32, 'inverse None')


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['zero', 'symmetric', 'periodization'])
def test_dwt1d_gradients_and_large(mode):
    torch.manual_seed(3)
    f = pw.DWT1DForward(J=3, wave='db4', mode=mode).cuda()
    i = pw.DWT1DInverse(wave='db4', mode=mode).cuda()
    x = torch.randn(3, 5, 4096, device='cuda', requires_grad=True)
    yl, yh = f(x)
    y = i((yl, yh))
    assert (y[..., :4096] - x).abs().max().item() < 1e-4                        # perfect reconstruction
    # <f(x), c> == <x, f^T(c)>: the backward pass is the reference's (synthesis with the analysis filters); it is the
    # true adjoint for zero padding and periodization
    gl, gh = torch.randn_like(yl), [torch.randn_like(h) for h in yh]
    lhs = (yl * gl).sum() + sum((h * g).sum() for h, g in zip(yh, gh))
    lhs.backward()
    if mode != 'symmetric':
        x2 = torch.randn_like(x)
        with torch.no_grad():
            zl, zh = f(x2)
            rhs = (zl * gl).sum() + sum((h * g).sum() for h, g in zip(zh, gh))
            assert abs(float((x2 * x.grad).sum()) - float(rhs)) <= 1e-3 * max(1.0, abs(float(rhs)))
    assert torch.isfinite(x.grad).all()
