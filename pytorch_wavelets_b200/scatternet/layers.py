"""``ScatLayer`` / ``ScatLayerj2``: DTCWT scattering layers (drop-in for the reference
``pytorch_wavelets/scatternet/layers.py:11-172``).  The default first-order layer (two-filter bank, no colour
combining) is ONE fused kernel; the variants (``combine_colour``, the 3-filter ``*_bp`` banks, ``ScatLayerj2``) run
their filter banks on the same kernels and compose the pointwise epilogue in ``scatternet/variants.py``."""
import torch
import torch.nn as nn

from pytorch_wavelets_b200.dtcwt.coeffs import biort as _biort, qshift as _qshift
from pytorch_wavelets_b200.dtcwt.lowlevel import prep_filt
from pytorch_wavelets_b200.scatternet import variants
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
        names = ('h0o', 'h1o', 'h2o') if biort == 'near_sym_b_bp' else ('h0o', 'h1o')
        self.bandpass_diag = biort == 'near_sym_b_bp'
        filts = _biort(biort)
        for n, arr in zip(names, filts[0::2]):      # h0o, g0o, h1o, g1o[, h2o, g2o] -> the analysis filters
            setattr(self, n, torch.nn.Parameter(prep_filt(arr, 1), False))
        self._names = names

    def forward(self, x):
        _, ch, r, c = x.shape
        if r % 2 != 0:
            x = torch.cat((x, x[:, :, -1:]), dim=2)
        if c % 2 != 0:
            x = torch.cat((x, x[:, :, :, -1:]), dim=3)
        if self.combine_colour:
            assert ch == 3
        if self.bandpass_diag or self.combine_colour:
            f = {n: getattr(self, n) for n in self._names}
            Z = variants.scat_j1(variants.KernelOps, x, f, self.mode, self.magbias, self.combine_colour)
        else:
            Z = ScatLayerj1_f.apply(x, self.h0o, self.h1o, self.mode, self.magbias, self.combine_colour)
        if not self.combine_colour:
            b, _, c, h, w = Z.shape
            Z = Z.reshape(b, 7 * c, h, w)
        return Z

    def extra_repr(self):
        return "biort='{}', mode='{}', magbias={}".format(self.biort, self.mode_str, self.magbias)


class ScatLayerj2(nn.Module):
    """Second-order scattering over two scales with the proper level-1 / level-2 DTCWT filters (reference
    ``layers.py:82-172``): (N, C, H, W) -> (N, 49*C, H/4, W/4), or (N, 51, H/4, W/4) with ``combine_colour``.
    The input is extended to a multiple of 8 by repeating its first / last rows and columns like the reference."""

    def __init__(self, biort='near_sym_a', qshift='qshift_a', mode='symmetric', magbias=1e-2, combine_colour=False):
        super().__init__()
        self.biort = biort
        self.qshift = biort          # (sic) the reference stores biort here, layers.py:105
        self.mode_str = mode
        self.mode = mode_to_int(mode)
        self.magbias = magbias
        self.combine_colour = combine_colour
        self.bandpass_diag = biort == 'near_sym_b_bp'
        if self.bandpass_diag:
            assert qshift == 'qshift_b_bp'
            names = ('h0o', 'h1o', 'h2o')
            qnames = ('h0a', 'h0b', 'h1a', 'h1b', 'h2a', 'h2b')
        else:
            names = ('h0o', 'h1o')
            qnames = ('h0a', 'h0b', 'h1a', 'h1b')
        for n, arr in zip(names, _biort(biort)[0::2]):
            setattr(self, n, torch.nn.Parameter(prep_filt(arr, 1), False))
        q = _qshift(qshift)              # h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b[, h2a, h2b, g2a, g2b]
        for n, arr in zip(qnames, [q[i] for i in (0, 1, 4, 5, 8, 9)[:len(qnames)]]):
            setattr(self, n, torch.nn.Parameter(prep_filt(arr, 1), False))
        self._names = names + qnames

    def forward(self, x):
        ch, r, c = x.shape[1:]
        rem = r % 8
        if rem != 0:
            rows_after, rows_before = (9 - rem) // 2, (8 - rem) // 2
            x = torch.cat((x[:, :, :rows_before], x, x[:, :, -rows_after:]), dim=2)
        rem = c % 8
        if rem != 0:
            cols_after, cols_before = (9 - rem) // 2, (8 - rem) // 2
            x = torch.cat((x[:, :, :, :cols_before], x, x[:, :, :, -cols_after:]), dim=3)
        if self.combine_colour:
            assert ch == 3
        if self.mode_str != 'symmetric':
            raise NotImplementedError()      # the reference's q-shift level only exists for symmetric extension
        f = {n: getattr(self, n) for n in self._names}
        Z = variants.scat_j2(variants.KernelOps, x, f, self.mode, self.magbias, self.combine_colour)
        if not self.combine_colour:
            b, _, c, h, w = Z.shape
            Z = Z.reshape(b, 49 * c, h, w)
        return Z

    def extra_repr(self):
        return "biort='{}', mode='{}', magbias={}".format(self.biort, self.mode_str, self.magbias)
