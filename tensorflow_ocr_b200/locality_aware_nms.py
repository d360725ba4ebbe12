"""Locality-aware NMS for the EAST decode — NOT in the reference (no ``lanms`` / ``nms_locality``
under the reference tree); same names as upstream argman/EAST
``locality_aware_nms.py``.  PARITY UNPINNED: checked against ``oracle/east.py`` only.
"""
from __future__ import annotations

import numpy as np
import torch

from . import head

__all__ = ["nms_locality", "nms_locality_batch"]


def nms_locality(polys, thres=0.3):
    """polys [N,9] (x0,y0,...,x3,y3,score) in row-major scan order -> survivors [M,9] (fp64),
    by descending score."""
    if len(polys) == 0:
        return np.zeros((0, 9))
    p, np_in = head.to_device(polys, torch.float64)
    offs = torch.tensor([0, p.shape[0]], dtype=torch.int32, device=p.device)
    out, n = head.lanms_raw(p.contiguous(), offs, thres)
    out = out[: int(n.item())]
    return out.cpu().numpy() if np_in else out


def nms_locality_batch(polys, offsets, thres=0.3):
    """Batched form: boxes of image b are rows offsets[b]:offsets[b+1].  Returns a list of arrays."""
    p, np_in = head.to_device(polys, torch.float64)
    offs, _ = head.to_device(np.asarray(offsets, np.int32) if not isinstance(offsets, torch.Tensor) else offsets,
                             torch.int32, p.device)
    out, n = head.lanms_raw(p.contiguous(), offs, thres)
    n = n.cpu().numpy()
    o = offs.cpu().numpy()
    res = [out[o[b]: o[b] + n[b]] for b in range(len(n))]
    return [r.cpu().numpy() for r in res] if np_in else res
