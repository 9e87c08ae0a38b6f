"""Stand-in for the PyWavelets package, used ONLY to import the reference in the build container.

The reference does ``import pywt`` (reference pytorch_wavelets/dwt/lowlevel.py:6,
dwt/transform2d.py:2, dwt/transform1d.py:2) but PyWavelets is not installed in this image and
there is no network.  On the hot path it needs exactly two things from it:
``pywt.Wavelet(name)`` (filter taps) and ``pywt.dwt_coeff_len`` (output length).  Both are
provided by ``pytorch_wavelets_b200.wavelets`` (published Daubechies construction / length rule).

TEST INFRASTRUCTURE: put on sys.path only by oracle/refshim.py.
"""
from pytorch_wavelets_b200.wavelets import Wavelet, dwt_coeff_len  # noqa: F401

__version__ = '0.0-standin'
