"""Mirror of the EAST geometry restore of the reference's ``datasets/icdar.py`` (lines 410-483)."""
from __future__ import annotations

from .. import head

__all__ = ["restore_rectangle_rbox", "restore_rectangle"]


def _float_dtype(x):
    """float64 inputs stay float64 (the reference computes in the dtype it is given); everything else
    (float32, float16, integers) is taken as float32."""
    import numpy as np
    import torch
    dt = x.dtype
    return torch.float64 if dt in (torch.float64, np.float64, np.dtype("float64")) else torch.float32


def restore_rectangle_rbox(origin, geometry, return_index=False):
    """datasets/icdar.py:410-479: ``origin`` [N,2], ``geometry`` [N,5] (top,right,bottom,left,theta)
    -> [N,4,2] float64.  Rows come out theta >= 0 first, then theta < 0 (icdar.py:479, quirk
    Q16); ``return_index=True`` (extension) also returns the input row of each output row."""
    o, np_in = head.to_device(origin, dtype=_float_dtype(origin))
    g, _ = head.to_device(geometry, dtype=_float_dtype(geometry), device=o.device)
    if o.dim() != 2 or o.shape[1] != 2 or g.dim() != 2 or g.shape[1] != 5 or g.shape[0] != o.shape[0]:
        raise ValueError("restore_rectangle expects origin [N,2] and geometry [N,5]")
    out, idx = head.restore_rectangle_raw(o, g, want_index=return_index)
    if np_in:
        out = out.cpu().numpy()
        idx = idx.cpu().numpy() if idx is not None else None
    return (out, idx) if return_index else out


def restore_rectangle(origin, geometry):
    """datasets/icdar.py:482-483."""
    return restore_rectangle_rbox(origin, geometry)
