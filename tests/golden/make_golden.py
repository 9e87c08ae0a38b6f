"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; imports it through oracle/refshim.py with
the pywt stand-in):

    python tests/golden/make_golden.py

Each fixture is an .npz with the seeded input, the reference module's stored filter buffers and
the reference's outputs, for one entry point of the hot path.  The GPU box has no /root/reference;
tests there (and the CPU tests of the oracle) read only these files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402

ref = refshim.load()


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: v.shape for k, v in out.items() if v.ndim})


def bufs(m, names):
    return {n: getattr(m, n).detach().numpy().ravel() for n in names}


def dwt_case(name, shape, J, wave, mode, seed):
    torch.manual_seed(seed)
    x = torch.randn(*shape)
    f = ref.DWTForward(J=J, wave=wave, mode=mode)
    i = ref.DWTInverse(wave=wave, mode=mode)
    yl, yh = f(x)
    y = i((yl, yh))
    yh_drop = list(yh)
    if J > 1:
        yh_drop[0] = None
    y_drop = i((yl, yh_drop))
    d = dict(x=x, yl=yl, y=y, y_drop0=y_drop, J=J, mode=mode, wave=wave)
    for j, h in enumerate(yh):
        d['yh%d' % j] = h
    d.update(bufs(f, ['h0_col', 'h1_col', 'h0_row', 'h1_row']))
    d.update(bufs(i, ['g0_col', 'g1_col', 'g0_row', 'g1_row']))
    save(name, **d)


def dtcwt_case(name, shape, J, biort, qshift, o_dim, ri_dim, mode, seed, skip_hps=False, scale=100.0):
    torch.manual_seed(seed)
    x = scale * torch.randn(*shape)
    f = ref.DTCWTForward(biort=biort, qshift=qshift, J=J, o_dim=o_dim, ri_dim=ri_dim, mode=mode,
                         skip_hps=skip_hps)
    i = ref.DTCWTInverse(biort=biort, qshift=qshift, o_dim=o_dim, ri_dim=ri_dim, mode=mode)
    yl, yh = f(x)
    yh_in = [None if h.shape == torch.Size([]) else h for h in yh]
    y = i((yl, yh_in))
    d = dict(x=x, yl=yl, y=y, J=J, mode=mode, biort=biort, qshift=qshift, o_dim=o_dim, ri_dim=ri_dim,
             skip=np.array([h is None for h in yh_in]))
    for j, h in enumerate(yh_in):
        if h is not None:
            d['yh%d' % j] = h
    d.update(bufs(f, ['h0o', 'h1o', 'h0a', 'h0b', 'h1a', 'h1b']))
    d.update(bufs(i, ['g0o', 'g1o', 'g0a', 'g0b', 'g1a', 'g1b']))
    save(name, **d)


def scat_case(name, shape, biort, mode, seed, magbias=1e-2):
    torch.manual_seed(seed)
    x = torch.randn(*shape)
    s = ref.ScatLayer(biort=biort, mode=mode, magbias=magbias)
    z = s(x)
    s2 = torch.nn.Sequential(ref.ScatLayer(biort=biort, mode=mode, magbias=magbias),
                             ref.ScatLayer(biort=biort, mode=mode, magbias=magbias))
    z2 = s2(x)
    save(name, x=x, z=z, z2=z2, h0o=s.h0o.detach().numpy().ravel(), h1o=s.h1o.detach().numpy().ravel(),
         mode=mode, biort=biort, magbias=magbias)


def scat_variant_case(name, shape, kind, seed, biort='near_sym_a', qshift='qshift_a', mode='symmetric',
                      combine_colour=False, magbias=1e-2):
    """ScatLayer variants of SURVEY 8(f)2: combine_colour, the 3-filter *_bp ("rot") filters, ScatLayerj2; the
    gradient with respect to the input is stored too (it exercises the reference's hand-written backward)."""
    torch.manual_seed(seed)
    x = torch.randn(*shape, requires_grad=True)
    if kind == 'j1':
        s = ref.ScatLayer(biort=biort, mode=mode, magbias=magbias, combine_colour=combine_colour)
    else:
        s = ref.ScatLayerj2(biort=biort, qshift=qshift, mode=mode, magbias=magbias, combine_colour=combine_colour)
    z = s(x)
    torch.manual_seed(seed + 1000)
    g = torch.randn_like(z)
    (z * g).sum().backward()
    save(name, x=x.detach(), z=z.detach(), g=g, dx=x.grad, kind=kind, biort=biort, qshift=qshift, mode=mode,
         combine_colour=int(combine_colour), magbias=magbias)


def dwt1d_case(name, shape, J, wave, mode, seed):
    torch.manual_seed(seed)
    x = torch.randn(*shape)
    f = ref.DWT1DForward(J=J, wave=wave, mode=mode)
    i = ref.DWT1DInverse(wave=wave, mode=mode)
    yl, yh = f(x)
    y = i((yl, yh))
    yh_drop = list(yh)
    yh_drop[0] = None
    y_drop = i((yl, yh_drop))
    d = dict(x=x, yl=yl, y=y, y_drop0=y_drop, J=J, mode=mode, wave=wave)
    for j, h in enumerate(yh):
        d['yh%d' % j] = h
    d.update(bufs(f, ['h0', 'h1']))
    d.update(bufs(i, ['g0', 'g1']))
    save(name, **d)


def prims_case(name, seed):
    """The reference's standalone 1-D primitives (dtcwt/lowlevel.py:70-239): odd AND even level-1 filter lengths (an even
    length gives N + 1 outputs, tests/test_colfilter.py:52-61), both extension modes, both q-shift phase-table parities
    (qshift_a: m/2 = 5 odd, qshift_b: m/2 = 7 odd, qshift_06: m/2 = 5, qshift_c: m/2 = 8 even), low- and high-pass."""
    from pytorch_wavelets.dtcwt import lowlevel as ll
    from pytorch_wavelets.dtcwt.coeffs import biort as _biort, qshift as _qshift
    torch.manual_seed(seed)
    x = torch.randn(2, 3, 24, 28)
    d = dict(x=x)
    rng = np.random.RandomState(seed)
    for L in (5, 7, 4, 6, 2):
        h = rng.randn(L)
        hp = ll.prep_filt(h, 1)
        d['filt%d_h' % L] = hp.numpy().ravel()
        for mode in ('symmetric', 'zero'):
            d['filt%d_col_%s' % (L, mode)] = ll.colfilter(x, hp, mode=mode)
            d['filt%d_row_%s' % (L, mode)] = ll.rowfilter(x, hp, mode=mode)
    for q in ('qshift_a', 'qshift_b', 'qshift_c'):
        h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b = _qshift(q)
        ha, hb = ll.prep_filt(h0a, 1), ll.prep_filt(h0b, 1)
        d['%s_ha' % q] = ha.numpy().ravel(); d['%s_hb' % q] = hb.numpy().ravel()
        for hp_ in (False, True):
            d['%s_coldfilt_%d' % (q, hp_)] = ll.coldfilt(x, ha, hb, highpass=hp_)
            d['%s_rowdfilt_%d' % (q, hp_)] = ll.rowdfilt(x, ha, hb, highpass=hp_)
            d['%s_colifilt_%d' % (q, hp_)] = ll.colifilt(x, ha, hb, highpass=hp_)
            d['%s_rowifilt_%d' % (q, hp_)] = ll.rowifilt(x, ha, hb, highpass=hp_)
    save(name, **d)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'prims':
    prims_case('prims_24x28', 95)
    sys.exit(0)

if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'extra':
    dtcwt_case('dtcwt_b_32_J3_96x64', (1, 1, 96, 64), 3, 'near_sym_b', 'qshift_32', 2, -1, 'symmetric', 90)
    sys.exit(0)

if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'dwt1d':
    k = 80
    for mode in ['zero', 'symmetric', 'reflect', 'periodic', 'periodization']:
        dwt1d_case('dwt1d_db4_%s_J3_200' % mode, (2, 3, 200), 3, 'db4', mode, k); k += 1
        dwt1d_case('dwt1d_db3_%s_J2_77' % mode, (1, 2, 77), 2, 'db3', mode, k); k += 1
    dwt1d_case('dwt1d_db1_zero_J4_1000', (1, 1, 1000), 4, 'db1', 'zero', k); k += 1
    dwt1d_case('dwt1d_db8_symmetric_J2_513', (2, 1, 513), 2, 'db8', 'symmetric', k); k += 1
    sys.exit(0)

if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'scatvar':
    scat_variant_case('scatv_j1_cc_32', (2, 3, 32, 32), 'j1', 60, combine_colour=True)
    scat_variant_case('scatv_j1_bp_32', (1, 2, 32, 32), 'j1', 61, biort='near_sym_b_bp')
    scat_variant_case('scatv_j1_bp_cc_zero_30x32', (1, 3, 30, 32), 'j1', 62, biort='near_sym_b_bp', mode='zero',
                      combine_colour=True)
    scat_variant_case('scatv_j2_a_32', (2, 2, 32, 32), 'j2', 63)
    scat_variant_case('scatv_j2_a_cc_40x36', (1, 3, 40, 36), 'j2', 64, combine_colour=True)
    scat_variant_case('scatv_j2_bp_32', (1, 2, 32, 32), 'j2', 65, biort='near_sym_b_bp', qshift='qshift_b_bp')
    scat_variant_case('scatv_j2_bp_cc_32', (1, 3, 32, 32), 'j2', 66, biort='near_sym_b_bp', qshift='qshift_b_bp',
                      combine_colour=True)
    scat_variant_case('scatv_j2_b_34x44', (1, 1, 34, 44), 'j2', 67, biort='near_sym_b', qshift='qshift_b')
    # (ScatLayerj2 with mode='zero' raises NotImplementedError in the reference: rowdfilt, dtcwt/lowlevel.py:142)
    sys.exit(0)

if __name__ == '__main__':
    # BASELINE.json config 1: DWTForward J=1 db4 zero on randn(4,3,64,64) -- the bit-check case
    dwt_case('dwt_c1_db4_zero_J1', (4, 3, 64, 64), 1, 'db4', 'zero', 0)
    k = 1
    for mode in ['zero', 'symmetric', 'reflect', 'periodic', 'periodization']:
        dwt_case('dwt_db4_%s_J3_64' % mode, (1, 2, 64, 64), 3, 'db4', mode, k); k += 1
        dwt_case('dwt_db3_%s_J2_37x50' % mode, (2, 1, 37, 50), 2, 'db3', mode, k); k += 1
    dwt_case('dwt_db1_symmetric_J3_33x64', (1, 2, 33, 64), 3, 'db1', 'symmetric', k); k += 1
    dwt_case('dwt_db8_zero_J2_96', (1, 1, 96, 96), 2, 'db8', 'zero', k); k += 1
    dwt_case('dwt_db2_symmetric_J2_127x126', (1, 1, 127, 126), 2, 'db2', 'symmetric', k); k += 1

    dtcwt_case('dtcwt_a_a_J3_64', (1, 2, 64, 64), 3, 'near_sym_a', 'qshift_a', 2, -1, 'symmetric', 20)
    dtcwt_case('dtcwt_a_a_J3_64_zero', (1, 2, 64, 64), 3, 'near_sym_a', 'qshift_a', 2, -1, 'zero', 21)
    dtcwt_case('dtcwt_a_a_J3_100', (1, 1, 100, 100), 3, 'near_sym_a', 'qshift_a', 2, -1, 'symmetric', 22)
    dtcwt_case('dtcwt_a_a_J4_99x100', (1, 1, 99, 100), 4, 'near_sym_a', 'qshift_a', 2, -1, 'symmetric', 23)
    dtcwt_case('dtcwt_a_a_J2_104x101', (1, 2, 104, 101), 2, 'near_sym_a', 'qshift_a', 2, -1, 'symmetric', 24)
    dtcwt_case('dtcwt_b_b_J3_72x56', (1, 1, 72, 56), 3, 'near_sym_b', 'qshift_b', 2, -1, 'symmetric', 25)
    dtcwt_case('dtcwt_ant_c_J3_64', (1, 1, 64, 64), 3, 'antonini', 'qshift_c', 2, -1, 'symmetric', 26)
    dtcwt_case('dtcwt_leg_d_J2_48', (1, 1, 48, 48), 2, 'legall', 'qshift_d', 2, -1, 'symmetric', 27)
    dtcwt_case('dtcwt_a_06_J2_40', (1, 1, 40, 40), 2, 'near_sym_a', 'qshift_06', 2, -1, 'symmetric', 28)
    for n, (o, r) in enumerate([(1, 2), (4, 5), (3, 1), (5, 2), (2, 3), (1, -1)]):
        dtcwt_case('dtcwt_a_a_J2_24x28_o%d_r%d' % (o, r % 6), (1, 2, 24, 28), 2, 'near_sym_a', 'qshift_a', o, r,
                   'symmetric', 30 + n)
    dtcwt_case('dtcwt_a_a_J3_64_skip1', (1, 2, 64, 64), 3, 'near_sym_a', 'qshift_a', 2, -1, 'symmetric', 40,
               skip_hps=[True, False, False])
    dtcwt_case('dtcwt_a_a_J3_64_skip2', (1, 2, 64, 64), 3, 'near_sym_a', 'qshift_a', 2, -1, 'symmetric', 41,
               skip_hps=[False, True, False])

    scat_case('scat_a_sym_32', (2, 3, 32, 32), 'near_sym_a', 'symmetric', 50)
    scat_case('scat_a_zero_31x29', (1, 2, 31, 29), 'near_sym_a', 'zero', 51)
    scat_case('scat_b_sym_32', (1, 1, 32, 32), 'near_sym_b', 'symmetric', 52)
