"""Tiny driver for ncu captures: a few forward passes of one transform (no timing here)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw

which = sys.argv[1] if len(sys.argv) > 1 else 'dwt'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = 'cuda'
torch.manual_seed(0)
if which == 'dwt':
    x = torch.randn(n, 32, 512, 512, device=dev)
    f = pw.DWTForward(J=3, wave='db4', mode='symmetric').to(dev)
elif which == 'dwtinv':
    x = torch.randn(n, 32, 512, 512, device=dev)
    c = pw.DWTForward(J=3, wave='db4', mode='symmetric').to(dev)(x)
    g = pw.DWTInverse(wave='db4', mode='symmetric').to(dev)
    f = lambda _: g(c)
elif which == 'dtcwt':
    x = torch.randn(n, 3, 1024, 1024, device=dev)
    f = pw.DTCWTForward(J=3).to(dev)
elif which == 'dtcwtinv':
    x = torch.randn(n, 3, 1024, 1024, device=dev)
    c = pw.DTCWTForward(J=3).to(dev)(x)
    g = pw.DTCWTInverse().to(dev)
    f = lambda _: g(c)
elif which == 'c5':
    x = torch.randn(n, 16, 2048, 2048, device=dev)
    f = pw.DWTForward(J=4, wave='db8', mode='zero').to(dev)
elif which == 'c5inv':
    x = torch.randn(n, 16, 2048, 2048, device=dev)
    c = pw.DWTForward(J=4, wave='db8', mode='zero').to(dev)(x)
    g = pw.DWTInverse(wave='db8', mode='zero').to(dev)
    f = lambda _: g(c)
elif which == 'scat':
    x = torch.randn(n, 3, 256, 256, device=dev)
    f = torch.nn.Sequential(pw.ScatLayer(), pw.ScatLayer()).to(dev)
with torch.no_grad():
    for _ in range(reps):
        y = f(x)
torch.cuda.synchronize()
print('done', which, n)
