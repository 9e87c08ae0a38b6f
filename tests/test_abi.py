"""CPU: libb200wave.so builds/loads and exports every symbol include/b200wave.h declares; host-side
helpers (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

from pytorch_wavelets_b200 import _build, _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    _build.build()
    return _ffi.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'b200wave.h')).read()
    declared = sorted(set(re.findall(r'\b(b200w_[a-z0-9_]+)\s*\(', hdr)))
    assert declared, 'no declarations found'
    for name in declared:
        assert hasattr(lib, name), 'libb200wave.so does not export %s' % name
    assert sorted(_ffi.SYMBOLS) == declared


def test_version_and_strerror(lib):
    assert lib.b200w_version() == 100
    assert lib.b200w_strerror(0) == b'ok'
    assert lib.b200w_strerror(-1) == b'Unkown pad type'
    assert lib.b200w_last_cuda_error() == b''


def test_length_rules(lib):
    # pywt.dwt_coeff_len / reference dwt/lowlevel.py:153 and :242-267
    assert lib.b200w_dwt_coeff_len(512, 8, 1) == 259
    assert lib.b200w_dwt_coeff_len(259, 8, 1) == 133
    assert lib.b200w_dwt_coeff_len(133, 8, 0) == 70
    assert lib.b200w_dwt_coeff_len(127, 8, 2) == 64
    assert lib.b200w_dwt_coeff_len(2048, 16, 0) == 1031
    assert lib.b200w_dwt_rec_len(259, 8, 1) == 512
    assert lib.b200w_dwt_rec_len(64, 8, 2) == 128
    assert lib.b200w_dwt_coeff_len(0, 8, 1) < 0


def test_argument_validation_without_gpu(lib):
    """Validation happens before any CUDA call, so bad arguments are reported on a CPU-only box."""
    f = (ctypes.c_float * 8)(*([0.5] * 8))
    fp = ctypes.cast(f, ctypes.c_void_p)
    buf = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    rc = lib.b200w_dwt_afb2d(buf, 64, 8, buf, 49, 7, buf, 1, 8, 8, fp, fp, 8, fp, fp, 8, 3, None)
    assert rc == -1  # 'constant' is not a filter-bank mode (reference ValueError)
    rc = lib.b200w_dwt_afb2d(buf, 64, 8, buf, 49, 7, buf, 1, 8, 8, fp, fp, 8, fp, fp, 8, 99, None)
    assert rc == -1
    rc = lib.b200w_dwt_afb2d(None, 64, 8, buf, 49, 7, buf, 1, 8, 8, fp, fp, 8, fp, fp, 8, 1, None)
    assert rc == -3
    hs = (ctypes.c_longlong * 6)(1, 1, 1, 1, 1, 1)
    rc = lib.b200w_dtcwt_fwd_j2plus(buf, 36, 6, buf, 9, 3, buf, hs, 1, 1, 6, 6, fp, fp, fp, fp, 8, None)
    assert rc == -2  # rows/cols must be a multiple of 4 (reference ValueError)
    rc = lib.b200w_dtcwt_fwd_j1(buf, 64, 8, buf, 64, 8, buf, hs, 1, 1, 8, 8, fp, 8, fp, 7, 1, None)
    assert rc == -4  # even-length level-1 filter


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_ffi, '_lib', None)
    monkeypatch.setattr(_ffi, 'SO_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_ffi.B200WaveError):
        _ffi.lib()
