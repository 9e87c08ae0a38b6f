"""``DWTForward`` / ``DWTInverse`` with the reference's constructor signatures, buffer names and return
structure (reference ``pytorch_wavelets/dwt/transform2d.py:7-148``), running on the B200 engine."""
import torch
import torch.nn as nn

from pytorch_wavelets_b200 import wavelets
from pytorch_wavelets_b200.dwt import lowlevel


def _resolve_wave(wave, analysis):
    """(col_lo, col_hi, row_lo, row_hi) filter arrays from a name, a Wavelet object or tuples of
    arrays -- reference transform2d.py:22-33 (analysis) and :91-102 (synthesis)."""
    if isinstance(wave, str):
        wave = wavelets.Wavelet(wave)
    if hasattr(wave, 'dec_lo') and hasattr(wave, 'rec_lo'):
        if analysis:
            c0, c1 = wave.dec_lo, wave.dec_hi
        else:
            c0, c1 = wave.rec_lo, wave.rec_hi
        return c0, c1, c0, c1
    if len(wave) == 2:
        return wave[0], wave[1], wave[0], wave[1]
    if len(wave) == 4:
        return wave[0], wave[1], wave[2], wave[3]
    raise ValueError('wave must be a name, a Wavelet, or a tuple of 2 or 4 filter arrays')


class DWTForward(nn.Module):
    """2-D DWT forward decomposition of an image batch (drop-in for the reference ``DWTForward``).

    Args:
        J (int): number of levels.
        wave (str | Wavelet | tuple(ndarray)): wavelet name (``haar``/``dbN`` built in, other families
            through PyWavelets if installed), an object with ``dec_lo/dec_hi/rec_lo/rec_hi``, or filter
            arrays ``(h0, h1)`` / ``(h0_col, h1_col, h0_row, h1_row)``.
        mode (str): 'zero', 'symmetric', 'reflect', 'periodic' or 'periodization'.

    ``forward(x)`` with x (N, C, H, W) float32 on a CUDA device returns ``(yl, yh)``: ``yl`` the final
    low-pass (N, C, H', W') and ``yh`` a list of J tensors (N, C, 3, H'', W'') (LH, HL, HH), finest first.
    """

    def __init__(self, J=1, wave='db1', mode='zero'):
        super().__init__()
        h0_col, h1_col, h0_row, h1_row = _resolve_wave(wave, analysis=True)
        filts = lowlevel.prep_filt_afb2d(h0_col, h1_col, h0_row, h1_row)
        self.register_buffer('h0_col', filts[0])
        self.register_buffer('h1_col', filts[1])
        self.register_buffer('h0_row', filts[2])
        self.register_buffer('h1_row', filts[3])
        self.J = J
        self.mode = mode

    def forward(self, x):
        mode = lowlevel.mode_to_int(self.mode)
        if self.J < 1:
            return x, []
        # same argument order as reference transform2d.py:70-71: the *_col buffers land on the Function's
        # h*_row parameters and therefore filter along W; the *_row buffers filter along H.
        # The level loop of the reference (:68-74) runs inside ONE C-ABI call: a single fused kernel launch for
        # all J levels where the pyramid kernel applies, one launch per level otherwise.
        out = lowlevel.DWTPyramid.apply(x, self.h0_col, self.h1_col, self.h0_row, self.h1_row, mode, self.J)
        return out[0], list(out[1:])


class DWTInverse(nn.Module):
    """2-D DWT inverse reconstruction (drop-in for the reference ``DWTInverse``).

    ``forward((yl, yh))`` takes the output format of :class:`DWTForward`; any entry of ``yh`` may be
    ``None`` (treated as zeros)."""

    def __init__(self, wave='db1', mode='zero'):
        super().__init__()
        g0_col, g1_col, g0_row, g1_row = _resolve_wave(wave, analysis=False)
        filts = lowlevel.prep_filt_sfb2d(g0_col, g1_col, g0_row, g1_row)
        self.register_buffer('g0_col', filts[0])
        self.register_buffer('g1_col', filts[1])
        self.register_buffer('g0_row', filts[2])
        self.register_buffer('g1_row', filts[3])
        self.mode = mode

    def forward(self, coeffs):
        yl, yh = coeffs
        ll = yl
        mode = lowlevel.mode_to_int(self.mode)
        for h in yh[::-1]:
            if h is not None:
                # 'unpad' added dimensions (reference transform2d.py:142-145); a strided view, not a copy
                if ll.shape[-2] > h.shape[-2]:
                    ll = ll[..., :-1, :]
                if ll.shape[-1] > h.shape[-1]:
                    ll = ll[..., :-1]
            ll = lowlevel.SFB2D.apply(ll, h, self.g0_col, self.g1_col, self.g0_row, self.g1_row, mode)
        return ll
