"""Per-level CUDA-vs-oracle error table (diagnostic; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import oracle as orc
from pytorch_wavelets_b200.dwt import lowlevel as dl
from pytorch_wavelets_b200.dtcwt import transform_funcs as tf
from pytorch_wavelets_b200.scatternet import lowlevel as sl
from pytorch_wavelets_b200.dtcwt._tables import TABLES
from pytorch_wavelets_b200.wavelets import Wavelet

dev = 'cuda'
def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    if a.shape != b.shape: return 'SHAPE %s vs %s' % (a.shape, b.shape)
    return '%.2e' % (np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
def rev(t, k): return np.array(TABLES[t][k])[::-1].copy()
rng = np.random.default_rng(0)
w = Wavelet('db4'); h0 = np.array(w.dec_lo[::-1]); h1 = np.array(w.dec_hi[::-1]); g0 = np.array(w.rec_lo); g1 = np.array(w.rec_hi)
for shape in [(1,2,64,64),(1,2,104,102),(2,3,37,50),(1,1,259,259),(1,2,40,72)]:
    x = rng.standard_normal(shape).astype(np.float32)
    for mode in (0, 1, 2, 4, 6):
        ll, hi = dl.afb2d_level(torch.from_numpy(x).to(dev), h0, h1, h0, h1, mode)
        oll, ohi = orc.dwt_afb2d(x, h0, h1, h0, h1, mode)
        y = dl.sfb2d_level(torch.from_numpy(oll).to(dev), torch.from_numpy(ohi).to(dev), g0, g1, g0, g1, mode)
        oy = orc.dwt_sfb2d(oll, ohi, g0, g1, g0, g1, mode)
        print('afb', shape, mode, rel(ll, oll), rel(hi, ohi), ' sfb', rel(y, oy))
h0o, h1o = rev('near_sym_a', 'h0o'), rev('near_sym_a', 'h1o')
g0o, g1o = rev('near_sym_a', 'g0o'), rev('near_sym_a', 'g1o')
q = [rev('qshift_a', k) for k in ('h0a', 'h1a', 'h0b', 'h1b')]
gq = [rev('qshift_a', k) for k in ('g0a', 'g1a', 'g0b', 'g1b')]
for shape in [(1,2,64,64),(1,2,104,102),(1,2,104,104),(2,3,38,70),(1,2,40,72),(1,1,32,32),(1,1,32,36),(1,1,36,32),(1,1,64,68),(1,1,16,16),(1,1,8,8)]:
    x = (100*rng.standard_normal(shape)).astype(np.float32)
    xt = torch.from_numpy(x).to(dev)
    for mode in (1, 0):
        ll, hi = tf.fwd_j1(xt, h0o, h1o, False, 2, 5, mode)
        oll, ohi = orc.dtcwt_fwd_j1(x, h0o, h1o, False, 2, -1, 'symmetric' if mode else 'zero')
        y = tf.inv_j1(torch.from_numpy(oll).to(dev), torch.from_numpy(ohi).to(dev), g0o, g1o, 2, 5, mode)
        oy = orc.dtcwt_inv_j1(oll, ohi, g0o, g1o, 2, -1, 'symmetric' if mode else 'zero')
        print('fwd_j1', shape, mode, rel(ll, oll), rel(hi, ohi), ' inv_j1', rel(y, oy))
    if shape[2] % 4 == 0 and shape[3] % 4 == 0:
        ll, hi = tf.fwd_j2plus(xt, *q, False, 2, 5)
        oll, ohi = orc.dtcwt_fwd_j2plus(x, *q)
        y = tf.inv_j2plus(torch.from_numpy(oll).to(dev), torch.from_numpy(ohi).to(dev), *gq, 2, 5)
        oy = orc.dtcwt_inv_j2plus(oll, ohi, *gq)
        print('fwd_j2plus', shape, rel(ll, oll), rel(hi, ohi), ' inv_j2plus', rel(y, oy))
    z, _, _ = sl.scat_j1(xt, h0o, h1o, 1, 1e-2, False)
    print('scat', shape, rel(z, orc.scat_j1(x, h0o, h1o, 'symmetric', 1e-2)))
torch.cuda.synchronize()
