"""Mirror of the reference's ``tool/bboxes.py`` evaluation functions (lines 158-282): the mask-raster Jaccard of
quadrilaterals and the greedy detection / ground-truth matching, on the GPU (``csrc/evalbox.cu``).

Same names and argument meaning as the reference; numpy in -> numpy out, CUDA tensors in -> CUDA tensors out."""
from __future__ import annotations

import numpy as np
import torch

from .. import head

__all__ = ["np_bboxes_jaccard", "bboxes_jaccard", "bboxes_matching", "bboxes_matching_batch"]


def _quads(bboxes, name):
    t, np_in = head.to_device(bboxes, dtype=torch.int32)
    if t.numel() % 8:
        raise ValueError("%s must hold 8 coordinates per box, got shape %s" % (name, tuple(t.shape)))
    return t.reshape(-1, 4, 2), np_in


def _gt_quads(gxs, gys, device=None):
    gx, np_in = head.to_device(gxs, dtype=torch.int32, device=device)
    gy, _ = head.to_device(gys, dtype=torch.int32, device=gx.device)
    if gx.dim() != 2 or gx.shape[1] != 4 or gx.shape != gy.shape:
        raise ValueError("gxs / gys must be [G,4]")
    if gx.shape[0] == 0:
        raise ValueError("no ground-truth boxes (the reference takes np.max of an empty array here)")
    return torch.stack([gx, gy], dim=-1).contiguous(), np_in


def _check_range(*arrays):
    for a in arrays:
        if isinstance(a, np.ndarray) and a.size and np.abs(a).max() >= (1 << 20):
            raise ValueError("box coordinates must be smaller than 2^20 in magnitude")


def np_bboxes_jaccard(bbox, gxs, gys):
    """tool/bboxes.py:252-282: Jaccard of the quadrilateral ``bbox`` (8,) against the G quadrilaterals
    (``gxs``, ``gys`` [G,4]), both rasterised like ``cv2.drawContours(thickness=-1)`` on the reference's mask (origin
    (0, 0), 10 pixels beyond the largest coordinate: negative coordinates are clipped as cv2 clips them).  float32 [G]."""
    _check_range(np.asarray(bbox) if not torch.is_tensor(bbox) else None,
                        np.asarray(gxs) if not torch.is_tensor(gxs) else None,
                        np.asarray(gys) if not torch.is_tensor(gys) else None)
    det, np_in = _quads(bbox, "bbox")
    if det.shape[0] != 1:
        raise ValueError("bbox must be one box of 8 coordinates")
    gts, _ = _gt_quads(gxs, gys, det.device)
    G = gts.shape[0]
    out = head.bboxes_matching_raw(det, gts, [1], [G], torch.zeros((G,), dtype=torch.uint8, device=det.device))
    return out["jaccard"].cpu().numpy() if np_in else out["jaccard"]


def bboxes_jaccard(bbox, gxs, gys):
    """tool/bboxes.py:247-250 wraps np_bboxes_jaccard in tf.py_func; here the op is native."""
    return np_bboxes_jaccard(bbox, gxs, gys)


def bboxes_matching(bboxes, gxs, gys, gignored, matching_threshold=0.5, scope=None):
    """tool/bboxes.py:158-246.  ``bboxes`` [N,8] detections in score order, ``gxs`` / ``gys`` [G,4], ``gignored``
    [G].  -> (n_gbboxes, tp_match [N] bool, fp_match [N] bool)."""
    _check_range(np.asarray(bboxes) if not torch.is_tensor(bboxes) else None,
                        np.asarray(gxs) if not torch.is_tensor(gxs) else None,
                        np.asarray(gys) if not torch.is_tensor(gys) else None)
    dets, np_in = _quads(bboxes, "bboxes")
    gts, _ = _gt_quads(gxs, gys, dets.device)
    gi, _ = head.to_device(np.asarray(gignored).astype(bool).astype(np.uint8) if not torch.is_tensor(gignored) else
                           (gignored != 0).to(torch.uint8), dtype=torch.uint8, device=dets.device)
    if gi.numel() != gts.shape[0]:
        raise ValueError("gignored must have one entry per ground-truth box")
    out = head.bboxes_matching_raw(dets, gts, [dets.shape[0]], [gts.shape[0]], gi.reshape(-1), matching_threshold)
    tp, fp, n = out["tp"].bool(), out["fp"].bool(), out["n_gbboxes"][0]
    if np_in:
        return np.int64(n.item()), tp.cpu().numpy(), fp.cpu().numpy()
    return n.to(torch.int64), tp, fp


def bboxes_matching_batch(bboxes, gxs, gys, gignored, matching_threshold=0.5):
    """The commented-out batch form (tool/bboxes.py:143-156) as one launch pair: sequences of per-image arrays in,
    (n_gbboxes [B], list of tp, list of fp) out."""
    B = len(bboxes)
    dets = [np.asarray(b, np.int32).reshape(-1, 4, 2) for b in bboxes]
    gts = [np.stack([np.asarray(x, np.int32), np.asarray(y, np.int32)], -1).reshape(-1, 4, 2) for x, y in zip(gxs, gys)]
    ign = [np.asarray(g).astype(bool).astype(np.uint8).reshape(-1) for g in gignored]
    _check_range(*dets, *gts)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    d, _ = head.to_device(np.concatenate(dets) if dets else np.zeros((0, 4, 2), np.int32), dtype=torch.int32, device=dev)
    g, _ = head.to_device(np.concatenate(gts), dtype=torch.int32, device=d.device)
    i, _ = head.to_device(np.concatenate(ign), dtype=torch.uint8, device=d.device)
    out = head.bboxes_matching_raw(d, g, [len(x) for x in dets], [len(x) for x in gts], i, matching_threshold)
    tp, fp = out["tp"].cpu().numpy().astype(bool), out["fp"].cpu().numpy().astype(bool)
    off = out["det_off"]
    return (out["n_gbboxes"].cpu().numpy().astype(np.int64), [tp[off[b]:off[b + 1]] for b in range(B)],
            [fp[off[b]:off[b + 1]] for b in range(B)])
