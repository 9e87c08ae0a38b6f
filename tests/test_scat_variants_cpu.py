"""CPU: the composition of the ScatLayer variants (scatternet/variants.py: combine_colour, 3-filter *_bp banks,
ScatLayerj2) against the reference's golden outputs, with the filter-bank primitives swapped for the oracle's
(the GPU tests run the same composition on the CUDA kernels).  Forward only here; the gradient goldens are checked
on the GPU, where the primitives are differentiable."""
import numpy as np
import pytest
import torch

import pytorch_wavelets_b200 as pw
from oracle import oracle as orc
from pytorch_wavelets_b200.scatternet import variants
from tests import util


class OracleOps(object):
    @staticmethod
    def _split(hi):
        hi = torch.from_numpy(hi)
        return hi[..., 0], hi[..., 1]

    @staticmethod
    def fwd_j1(x, h0, h1, mode):
        ll, hi = orc.dtcwt_fwd_j1(x.numpy(), h0.detach().numpy(), h1.detach().numpy(), False, 1, -1,
                                  'symmetric' if mode == 1 else 'zero')
        return (torch.from_numpy(ll),) + OracleOps._split(hi)

    @staticmethod
    def fwd_j2plus(x, h0a, h1a, h0b, h1b, mode):
        ll, hi = orc.dtcwt_fwd_j2plus(x.numpy(), *[f.detach().numpy() for f in (h0a, h1a, h0b, h1b)], False, 1, -1)
        return (torch.from_numpy(ll),) + OracleOps._split(hi)


@pytest.mark.parametrize('name', util.fixtures('scatv_'))
def test_variant_composition_matches_reference(name, monkeypatch):
    g = util.load(name)
    kw = dict(biort=str(g['biort']), mode=str(g['mode']), magbias=float(g['magbias']),
              combine_colour=bool(int(g['combine_colour'])))
    if str(g['kind']) == 'j2':
        m = pw.ScatLayerj2(qshift=str(g['qshift']), **kw)
    else:
        m = pw.ScatLayer(**kw)
    monkeypatch.setattr(variants, 'KernelOps', OracleOps)
    z = m(torch.from_numpy(g['x']))
    assert tuple(z.shape) == g['z'].shape
    util.assert_close(z.numpy(), g['z'], util.RTOL_F32, name)


def test_scatlayerj2_zero_mode_raises_like_the_reference():
    with pytest.raises(NotImplementedError):
        pw.ScatLayerj2(mode='zero')(torch.zeros(1, 1, 16, 16))
