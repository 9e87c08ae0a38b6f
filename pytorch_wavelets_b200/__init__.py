"""pytorch_wavelets_b200 -- B200 (sm_100a) engine for the 2-D wavelet filterbank hot path of
fbcotter/pytorch_wavelets, behind the reference's nn.Module API.

Same export list and aliases as the reference package (``pytorch_wavelets/__init__.py:1-36``) for the
classes on the hot path and its direct callers.
Every transform runs in hand-written CUDA kernels through the C ABI of ``libb200wave.so``; there is no
CPU or eager fallback.
"""
__all__ = [
    '__version__',
    'DTCWTForward',
    'DTCWTInverse',
    'DWTForward',
    'DWTInverse',
    'DTCWT',
    'IDTCWT',
    'DWT',
    'IDWT',
    'DWT2D',
    'IDWT2D',
    'DWT1DForward',
    'DWT1DInverse',
    'DWT1D',
    'IDWT1D',
    'ScatLayer',
    'ScatLayerj2',
]

from pytorch_wavelets_b200._version import __version__
from pytorch_wavelets_b200.dtcwt.transform2d import DTCWTForward, DTCWTInverse
from pytorch_wavelets_b200.dwt.transform1d import DWT1DForward, DWT1DInverse
from pytorch_wavelets_b200.dwt.transform2d import DWTForward, DWTInverse
from pytorch_wavelets_b200.scatternet import ScatLayer, ScatLayerj2

# aliases, as in the reference
DTCWT = DTCWTForward
IDTCWT = DTCWTInverse
DWT = DWTForward
IDWT = DWTInverse
DWT2D = DWT
IDWT2D = IDWT
DWT1D = DWT1DForward
IDWT1D = DWT1DInverse
