"""Filter preparation for the DTCWT modules (reference ``pytorch_wavelets/dtcwt/lowlevel.py:58-67``).
The 1-D primitives of that file (colfilter, coldfilt, colifilt, q2c, c2q ...) have no standalone
counterpart here: they are fused inside the per-level CUDA kernels."""
import numpy as np
import torch


def prep_filt(h, c, transpose=False):
    """Column filter -> reversed (c,1,L,1) tensor (or (c,1,1,L) with transpose), default dtype."""
    h = np.atleast_2d(np.asarray(h))
    if h.shape[0] == 1:
        h = h.T
    h = h[::-1]
    h = h[None, None, :]
    h = np.repeat(h, repeats=c, axis=0)
    if transpose:
        h = h.transpose((0, 1, 3, 2))
    h = np.copy(h)
    return torch.tensor(h, dtype=torch.get_default_dtype())
