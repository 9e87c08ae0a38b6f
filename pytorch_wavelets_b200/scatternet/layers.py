"""``ScatLayer``: one order of DTCWT scattering at a single scale (drop-in for the reference
``pytorch_wavelets/scatternet/layers.py:11-79``, non-colour-combining, two-filter biorthogonal path)."""
import torch
import torch.nn as nn

from pytorch_wavelets_b200.dtcwt.coeffs import biort as _biort
from pytorch_wavelets_b200.dtcwt.lowlevel import prep_filt
from pytorch_wavelets_b200.scatternet.lowlevel import ScatLayerj1_f, mode_to_int


class ScatLayer(nn.Module):
    """First-order scattering layer: level-1 DTCWT, complex magnitude (smoothed by ``magbias``) of the six
    orientations and a 2x2 average of the low-pass, stacked on the channel dimension.

    Input (N, C, H, W) -> output (N, 7*C, H/2, W/2): the first C channels are the low-pass, the next 6C
    the magnitudes (orientation-major).  Stack two layers for a second-order scatternet.
    """

    def __init__(self, biort='near_sym_a', mode='symmetric', magbias=1e-2, combine_colour=False):
        super().__init__()
        self.biort = biort
        self.mode_str = mode
        self.mode = mode_to_int(mode)
        self.magbias = magbias
        self.combine_colour = combine_colour
        if biort == 'near_sym_b_bp':
            raise NotImplementedError("biort='near_sym_b_bp' (3-filter rotationally symmetric variant) is outside "
                                      "the accelerated hot path (SURVEY 8(f) rank 2)")
        if combine_colour:
            raise NotImplementedError('combine_colour=True is outside the accelerated hot path (SURVEY 8(f) rank 2)')
        self.bandpass_diag = False
        h0o, _, h1o, _ = _biort(biort)[:4]
        self.h0o = torch.nn.Parameter(prep_filt(h0o, 1), False)
        self.h1o = torch.nn.Parameter(prep_filt(h1o, 1), False)

    def forward(self, x):
        _, ch, r, c = x.shape
        if r % 2 != 0:
            x = torch.cat((x, x[:, :, -1:]), dim=2)
        if c % 2 != 0:
            x = torch.cat((x, x[:, :, :, -1:]), dim=3)
        Z = ScatLayerj1_f.apply(x, self.h0o, self.h1o, self.mode, self.magbias, self.combine_colour)
        b, _, c, h, w = Z.shape
        return Z.view(b, 7 * c, h, w)

    def extra_repr(self):
        return "biort='{}', mode='{}', magbias={}".format(self.biort, self.mode_str, self.magbias)
