"""Mirror of the reference's ``tool/pixellink_fn.py`` decode helper (lines 120-158)."""
from __future__ import annotations

import numpy as np
import torch

from .. import head

__all__ = ["pixel_detect", "tf_pixel_detect"]


def pixel_detect(score_map, geo_map, score_map_thresh=0.8, link_thresh=0.8):
    """tool/pixellink_fn.py:120-154: ``score > thr_p`` AND, for all 8 directions,
    ``link_d[...,1] >= thr_l``  -> uint8 [H,W].

    ``score_map`` [1,H,W,1] probabilities, ``geo_map`` [8,1,H,W,2] link softmax outputs;
    like the reference only 4-D input and batch index 0 are handled (quirk Q9).
    """
    sm, np_in = head.to_device(score_map)
    if sm.dim() != 4:
        raise ValueError("pixel_detect handles 4-D score maps only (tool/pixellink_fn.py:131-133)")
    gm, _ = head.to_device(geo_map, device=sm.device)
    if gm.dim() != 5 or gm.shape[0] != 8 or gm.shape[-1] != 2:
        raise ValueError("geo_map must be [8,1,H,W,2]")
    score = sm[0, :, :, 0].contiguous()
    link = gm[:, 0].contiguous()
    out = head.pixel_detect_raw(score, link, score_map_thresh, link_thresh)
    return out.cpu().numpy() if np_in else out


def tf_pixel_detect(score_map, geo_map, score_map_thresh, link_thresh):
    """tool/pixellink_fn.py:156-158 wraps pixel_detect in tf.py_func; here the op is native."""
    return pixel_detect(score_map, geo_map, score_map_thresh, link_thresh)
