"""The ScatLayer variants of the reference beyond the fused two-filter first-order layer:
``combine_colour``, the three-filter rotationally symmetric banks (``near_sym_b_bp`` / ``qshift_b_bp``) and the
two-scale ``ScatLayerj2`` (reference ``scatternet/lowlevel.py:140-599``, ``dtcwt/transform_funcs.py:124-149,
187-223,252-276,310-340``).

Every filter bank runs on the CUDA kernels behind the C ABI (``FWD_J1`` / ``FWD_J2PLUS`` and, through autograd,
``INV_J1`` / ``INV_J2PLUS`` with the reference's swapped-tree backward filters); what is composed here is the
pointwise epilogue (smoothed magnitude, 2x2 mean, stacking).  The 3-filter banks are two launches of the 2-filter
kernel: the diagonal band ``hh = col_h2(row_h2(x))`` is exactly the ``hh`` of a bank built from ``(h0, h2)``, so
orientations 45 and 135 degrees are taken from a second pass -- the same arithmetic as the reference's
``fwd_j1_rot`` / ``fwd_j2plus_rot``, and autograd then yields the reference's ``inv_*_rot`` backward.

``ops`` is the set of bank primitives; the CPU tests inject oracle-backed ones to check this composition against
the reference's golden outputs without a GPU.
"""
import torch
import torch.nn.functional as F

from pytorch_wavelets_b200.dtcwt import transform_funcs as tf


class KernelOps(object):
    """Bank primitives on the B200 kernels; band-pass returned as (re, im) of shape (N, 6, C, h, w)."""

    @staticmethod
    def fwd_j1(x, h0, h1, mode):
        ll, hi = tf.FWD_J1.apply(x, h0, h1, False, 1, -1, mode)
        return ll, hi[..., 0], hi[..., 1]

    @staticmethod
    def fwd_j2plus(x, h0a, h1a, h0b, h1b, mode):
        ll, hi = tf.FWD_J2PLUS.apply(x, h0a, h1a, h0b, h1b, False, 1, -1, mode)
        return ll, hi[..., 0], hi[..., 1]


def _swap_diag(a, b):
    """Orientations 45 / 135 degrees (the hh band) from the second pass, the rest from the first."""
    return torch.stack((a[:, 0], b[:, 1], a[:, 2], a[:, 3], b[:, 4], a[:, 5]), dim=1)


def bank_j1(ops, x, f, mode):
    ll, re, im = ops.fwd_j1(x, f['h0o'], f['h1o'], mode)
    if 'h2o' in f:
        _, re2, im2 = ops.fwd_j1(x, f['h0o'], f['h2o'], mode)
        re, im = _swap_diag(re, re2), _swap_diag(im, im2)
    return ll, re, im


def bank_j2(ops, x, f, mode):
    ll, re, im = ops.fwd_j2plus(x, f['h0a'], f['h1a'], f['h0b'], f['h1b'], mode)
    if 'h2a' in f:
        _, re2, im2 = ops.fwd_j2plus(x, f['h0a'], f['h2a'], f['h0b'], f['h2b'], mode)
        re, im = _swap_diag(re, re2), _swap_diag(im, im2)
    return ll, re, im


def smooth_mag(re, im, bias, combine_colour):
    """sqrt(re^2 + im^2 + b^2) - b, summed over the colour channel (dim 2, kept) when ``combine_colour``."""
    e = re * re + im * im
    if combine_colour:
        e = e.sum(dim=2, keepdim=True)
    return torch.sqrt(e + bias * bias) - bias


def scat_j1(ops, x, f, mode, bias, combine_colour):
    """One scale (reference ScatLayerj1_f / ScatLayerj1_rot_f): (N,7,C,h,w), or (N,9,h,w) when combining colours."""
    assert x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
    ll, re, im = bank_j1(ops, x, f, mode)
    ll = F.avg_pool2d(ll, 2)
    r = smooth_mag(re, im, bias, combine_colour)
    if combine_colour:
        return torch.cat((ll, r[:, :, 0]), dim=1)
    return torch.cat((ll[:, None], r), dim=1)


def scat_j2(ops, x, f, mode, bias, combine_colour):
    """Two scales, second order (reference ScatLayerj2_f / ScatLayerj2_rot_f): (N,49,C,h,w) or (N,51,h,w)."""
    assert x.shape[2] % 8 == 0 and x.shape[3] % 8 == 0
    s0, re, im = bank_j1(ops, x, f, mode)
    s1_j1 = smooth_mag(re, im, bias, combine_colour)                    # (N,6,C|1,H/2,W/2)
    s0, re, im = bank_j2(ops, s0, f, mode)
    s1_j2 = smooth_mag(re, im, bias, combine_colour)                    # (N,6,C|1,H/4,W/4)
    s0 = F.avg_pool2d(s0, 2)
    p = s1_j1.shape
    u = s1_j1[:, :, 0] if combine_colour else s1_j1.reshape(p[0], 6 * p[2], p[3], p[4])
    u, re, im = bank_j1(ops, u, f, mode)                                # second order on the first-scale magnitudes
    s2 = smooth_mag(re, im, bias, False)
    q = s2.shape
    u = F.avg_pool2d(u, 2)
    if combine_colour:
        return torch.cat((s0, u, s1_j2[:, :, 0], s2.reshape(q[0], 36, q[3], q[4])), dim=1)
    s2 = s2.reshape(q[0], 36, q[2] // 6, q[3], q[4])
    u = u.reshape(p[0], 6, p[2], p[3] // 2, p[4] // 2)
    return torch.cat((s0[:, None], u, s1_j2, s2), dim=1)
