"""bench.py --impl aten times an independently written ATen (conv2d + gather) version of the two headline transforms;
this checks on the CPU that it is the same operator as the oracle (so the GPU comparison compares like with like)."""
import numpy as np
import torch

from oracle import oracle as orc
from pytorch_wavelets_b200 import wavelets
from pytorch_wavelets_b200.dtcwt import coeffs
from tools import aten_arm


def test_aten_dwt_matches_oracle():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 40, 64)).astype(np.float32)
    w = wavelets.Wavelet('db4')
    h0, h1 = np.array(w.dec_lo[::-1], np.float32), np.array(w.dec_hi[::-1], np.float32)
    oyl, oyh = orc.dwt_forward(x, (h0, h1, h0, h1), 3, 'symmetric')
    yl, yh = aten_arm.dwt_fwd(torch.from_numpy(x), torch.from_numpy(h0), torch.from_numpy(h1), 3)
    assert np.abs(yl.numpy() - oyl).max() < 1e-5
    for a, b in zip(yh, oyh):
        assert a.shape == b.shape and np.abs(a.numpy() - b).max() < 1e-5


def test_aten_dtcwt_matches_oracle():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 2, 64, 96)).astype(np.float32)
    h0o, _, h1o, _ = coeffs.biort('near_sym_a')
    q = coeffs.qshift('qshift_a')
    l1 = (h0o[::-1].ravel(), h1o[::-1].ravel())
    # stored (reversed) q-shift filters in the oracle's order (h0a, h0b, h1a, h1b): see bench.cpu_sample
    qs = tuple(q[i][::-1].ravel() for i in (0, 1, 4, 5))
    oyl, oyh = orc.dtcwt_forward(x, l1, qs, 3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    yl, yh = aten_arm.dtcwt_fwd(t(x), t(l1[0]), t(l1[1]), t(qs[0]), t(qs[1]), t(qs[2]), t(qs[3]), 3)
    assert np.abs(yl.numpy() - oyl).max() < 1e-4
    for a, b in zip(yh, oyh):
        assert a.shape == b.shape and np.abs(a.numpy() - b).max() < 1e-4
