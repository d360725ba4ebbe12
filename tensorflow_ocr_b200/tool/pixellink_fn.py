"""Mirror of the reference's ``tool/pixellink_fn.py``: the decode helper (lines 120-158) and the
ground-truth generator ``generate_rbox`` (lines 49-117)."""
from __future__ import annotations

import numpy as np
import torch

from .. import head

__all__ = ["pixel_detect", "tf_pixel_detect", "generate_rbox", "tf_pixellink_get_rbox", "link_labels_from_ids",
           "rasterize_polygons"]


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


def link_labels_from_ids(poly_mask):
    """The double loop of generate_rbox (tool/pixellink_fn.py:81-109 with valid_link :9-47) on the GPU.

    ``poly_mask`` uint8 [H,W] or [B,H,W] (0 background, i = polygon i) -> link map fp32 [...,8]."""
    t, np_in = head.to_device(poly_mask, dtype=torch.uint8)
    single = t.dim() == 2
    link, _ = head.link_labels_raw(t[None] if single else t, want_pixel=False)
    link = link[0] if single else link
    return link.cpu().numpy() if np_in else link


def _canvas_quads(h, w, xs, ys):
    """Vertices as the reference makes them (tool/pixellink_fn.py:68-76): float32 products, truncated to int32."""
    vx = np.asarray(xs, dtype=np.float32) * w      # float32 products, like xs[idx, :] * w
    vy = np.asarray(ys, dtype=np.float32) * h
    return np.stack([vx, vy], axis=-1).astype(np.int32).reshape(-1, 4, 2)


def _rasterize_device(h, w, xs, ys):
    h, w = int(h), int(w)
    quads = _canvas_quads(h, w, xs, ys)
    dev = head._require_gpu(None)
    out = head.fill_quads_raw(torch.as_tensor(quads).to(dev), [len(quads)], h, w, h // 4, w // 4, mode=1,
                              want=("last", "ids_u8", "score"))    # Python-2 integer division in the reference (:56-57)
    return out["score"][0], out["ids_u8"][0]


def rasterize_polygons(h, w, xs, ys):
    """tool/pixellink_fn.py:60-79: every polygon filled at full resolution — 1.0 into the score map, its 1-based
    index into the uint8 id map — then both shrunk to (w//4, h//4) with cv2.resize(INTER_NEAREST).  On the GPU
    (`plh_fill_quads`): bit-identical to the OpenCV calls, without drawing the full-resolution canvases.
    -> (res_score_map float32 [h//4, w//4], poly_mask uint8 [h//4, w//4])."""
    score, ids = _rasterize_device(h, w, xs, ys)
    return score.cpu().numpy(), ids.cpu().numpy()


def generate_rbox(h, w, xs, ys, bboxes, ignored):
    """tool/pixellink_fn.py:53-111.  Same arguments and returns (res_score_map [h/4,w/4] fp32,
    res_link_map [h/4,w/4,8] fp32, show_bboxes [200,4] fp32) as numpy arrays.

    Both halves run on the GPU: the polygon rasterisation + nearest shrink (`plh_fill_quads`) and the per-pixel
    link-label loop — O(pixels x 8) interpreted Python in the reference (`plh_link_labels`); the id map never
    leaves the device in between."""
    if len(xs) != len(ignored):
        raise AssertionError("the length of xs and ignored must be the same, but got %s and %s" % (len(xs), len(ignored)))
    n = len(xs)
    show_bboxes = np.zeros((200, 4), dtype=np.float32)   # fixed 200 rows (:65)
    show_bboxes[:n] = np.asarray(bboxes)[:n]
    score, ids = _rasterize_device(h, w, xs, ys)
    link, _ = head.link_labels_raw(ids[None], want_pixel=False)
    return score.cpu().numpy(), link[0].cpu().numpy(), show_bboxes


def tf_pixellink_get_rbox(img_size, xs, ys, bboxes, ignored):
    """tool/pixellink_fn.py:113-118 wraps generate_rbox in tf.py_func; here it is called directly."""
    h, w = img_size
    return generate_rbox(h, w, xs, ys, bboxes, ignored)
