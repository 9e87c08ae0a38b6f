"""``ScatLayerj1_f``: one first-order DTCWT scattering layer as a single fused kernel (level-1 DTCWT,
2x2 mean of the low-pass, smoothed magnitude of the six orientations, channel stacking) -- drop-in for
the reference Function of the same name (``pytorch_wavelets/scatternet/lowlevel.py:71-137``,
``combine_colour=False`` path)."""
import torch
import torch.nn.functional as F
from torch.autograd import Function

from pytorch_wavelets_b200 import _ffi
from pytorch_wavelets_b200.dtcwt.transform_funcs import inv_j1
from pytorch_wavelets_b200.dwt.lowlevel import int_to_mode, mode_to_int  # noqa: F401  (re-exported like the reference)


def scat_j1(x, h0o, h1o, mode, bias, want_aux):
    """z (N,7,C,H/2,W/2) [, re/r, im/r (N,6,C,H/2,W/2)] from x (N,C,H,W), H and W even."""
    dt = _ffi.require_cuda_real(x, 'x')
    L = _ffi.lib()
    h0, h1 = _ffi.host_taps(h0o), _ffi.host_taps(h1o)
    x = x.contiguous()
    N, C, H, W = x.shape
    z = x.new_empty((N, 7, C, H // 2, W // 2))
    dre = dim = None
    if want_aux:
        dre = x.new_empty((N, 6, C, H // 2, W // 2))
        dim = torch.empty_like(dre)
    if N * C > 0:
        with torch.cuda.device(x.device), _ffi.span('scat_j1 %dx%d' % (H, W),
                                                    N * C * H * W * (4 + 7 + (12 if want_aux else 0))):
            rc = _ffi.entry('b200w_scat_j1', dt)(x.data_ptr(), z.data_ptr(), None if dre is None else dre.data_ptr(),
                                 None if dim is None else dim.data_ptr(), N, C, H, W, h0.p(dt), h0.n, h1.p(dt), h1.n,
                                 mode, float(bias), _ffi.stream_of(x))
        _ffi.check(rc, 'b200w_scat_j1')
    return z, dre, dim


class ScatLayerj1_f(Function):
    """``apply(x, h0o, h1o, mode, bias, combine_colour)`` -> Z of shape (N, 7, C, H/2, W/2)."""

    @staticmethod
    def forward(ctx, x, h0o, h1o, mode, bias, combine_colour):
        if combine_colour:
            raise NotImplementedError('combine_colour=True is outside the accelerated hot path (SURVEY 8(f) rank 2)')
        ctx.in_shape = x.shape
        batch, ch, r, c = x.shape
        assert r % 2 == c % 2 == 0
        mode = int(mode)
        int_to_mode(mode)
        ctx.mode = mode
        want = bool(x.requires_grad)
        ctx.taps = (_ffi.host_taps(h0o), _ffi.host_taps(h1o))
        z, dre, dim = scat_j1(x, ctx.taps[0], ctx.taps[1], mode, bias, want)
        if want:
            ctx.save_for_backward(dre, dim)
        return z

    @staticmethod
    def backward(ctx, dZ):
        dX = None
        if ctx.needs_input_grad[0]:
            h0o, h1o = ctx.taps
            drdx, drdy = ctx.saved_tensors
            dYl, dr = dZ[:, 0], dZ[:, 1:]
            ll = 1 / 4 * F.interpolate(dYl, scale_factor=2, mode='nearest')
            # band-pass gradient written straight into the (n, o, c, h, w, re/im) layout: re/im adjacent and columns
            # contiguous is all the streaming level-1 synthesis kernel needs (the orientation / channel order travels as
            # strides), so the backward runs on the fast path (round 1: (r, n, o, c, h, w) -> the 4x slower generic kernel)
            if dr.requires_grad:      # (create_graph=True: `out=` is not differentiable)
                highs = torch.stack((dr * drdx, dr * drdy), dim=-1)
            else:
                highs = dr.new_empty(dr.shape + (2,))
                torch.mul(dr, drdx, out=highs[..., 0])
                torch.mul(dr, drdy, out=highs[..., 1])
            dX = _inv_j1_o1_ri_last(ll, highs, h0o, h1o, ctx.mode)
        return (dX,) + (None,) * 5


def _inv_j1_o1_ri_last(ll, highs, g0, g1, mode):
    """inv_j1 for a band-pass tensor laid out (N, 6, C, h, w, 2).  transform_funcs._layout(o5, ri) inserts 'o' into
    (n,c,h,w) at o5 then 'r' at ri: o5=1 -> (n,o,c,h,w); ri=5 -> (n,o,c,h,w,r)."""
    return inv_j1(ll, highs, g0, g1, 1, 5, mode)
