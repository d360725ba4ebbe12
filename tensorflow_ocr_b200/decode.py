"""PixelLink inference decode — the part the reference only has inline in its scripts
(test_pixellink_fast.py:95-217, test_pixellink.py:94-230, test.py:24-74,182-218).

``decode_pixellink`` is the explicitly named addition of SURVEY.md §8b; ``pixel_detect``,
``order_points`` and ``sort_poly`` keep the names they have in the reference's ``test.py``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import head
from .tool.pixellink_fn import pixel_detect as _pixel_detect_fn

__all__ = ["decode_pixellink", "pixel_detect", "contour_boxes", "cv_contour_order", "order_points", "sort_poly",
           "write_result_txt"]


def decode_pixellink(pixel_logits, link_logits, pixel_thresh=0.8, link_thresh=0.9, min_size=10,
                     scale=(4.0, 3.75), max_boxes=128):
    """Logits -> (labels, boxes, counts), test_pixellink_fast.py:110-202 on the GPU.

    pixel_logits [B,H,W,2] (or [H,W,2]), link_logits [B,H,W,16].
      labels int32 [B,H,W]: -1 background / filtered component, else the component's
             minimum linear pixel index (canonical labelling)
      boxes  list of B int32 arrays [n_b,4,2] in cv2.boxPoints order, components in
             ascending label order (np.int0(cv2.boxPoints(cv2.minAreaRect(pts))))
      counts list of B int32 arrays [n_b]: pixel count of each component
    numpy in -> numpy out; CUDA tensors in -> CUDA tensors (one sync to read n_boxes).

    Deviation from the script (quirk Q10): the reference groups pixels by a DIRECTED depth-first search whose
    result depends on Python-2 dict order when link decisions are asymmetric; this function returns the
    weakly-connected components of the same edge set.  Identical whenever link decisions are symmetric
    (tests/golden/link_graph.npz pins that against the script's own lines); with asymmetric links every
    group of the script lies inside one component returned here, i.e. components are never split, only
    merged (tests/test_gpu_decode.py::test_decode_asymmetric_links_vs_literal_dfs quantifies it).
    """
    pl, np_in = head.to_device(pixel_logits)
    ll, _ = head.to_device(link_logits, device=pl.device)
    single = pl.dim() == 3
    if single:
        pl, ll = pl.unsqueeze(0), ll.unsqueeze(0)
    cfg = head.DecodeConfig(pixel_thresh, link_thresh, min_size, (float(scale[0]), float(scale[1])), max_boxes)
    out = head.decode_raw(pl, ll, cfg, want_rects=False)
    n = out["n_boxes"].cpu().numpy()
    if (n > max_boxes).any():
        raise ValueError("an image has %d components; raise max_boxes (=%d)" % (int(n.max()), max_boxes))
    labels = out["labels"]
    boxes = [out["boxes"][b, : n[b]] for b in range(len(n))]
    counts = [out["comp"][b, : n[b], 1] for b in range(len(n))]
    if np_in:
        labels = labels.cpu().numpy()
        boxes = [b.cpu().numpy() for b in boxes]
        counts = [c.cpu().numpy() for c in counts]
    if single:
        return labels[0], boxes[0], counts[0]
    return labels, boxes, counts


def pixel_detect(score_map, geo_map, score_map_thresh=0.8, link_thresh=0.8):
    """Name of test.py:45-74.  The reference's twin there is buggy (`res[link_text[0],
    link_text[1]] = 0` clears two cells instead of every failing pixel, and raises when
    a direction has < 2 failing pixels — SURVEY.md quirk Q8); this implements the intended
    semantics of tool/pixellink_fn.py:120-154 on test.py's input layout:
    ``score_map`` [1,H,W,1], ``geo_map`` [1,H,W,16] with channel 2i+1 = link_i score."""
    sm, np_in = head.to_device(score_map)
    gm, _ = head.to_device(geo_map, device=sm.device)
    if sm.dim() != 4 or gm.dim() != 4 or gm.shape[-1] != 16:
        raise ValueError("expected score_map [1,H,W,1] and geo_map [1,H,W,16]")
    H, W = sm.shape[1], sm.shape[2]
    g = gm[0].reshape(H, W, 8, 2).permute(2, 0, 1, 3).unsqueeze(1).contiguous()  # [8,1,H,W,2]
    out = _pixel_detect_fn(sm, g, score_map_thresh, link_thresh)
    return out.cpu().numpy() if np_in else out


def cv_contour_order(info):
    """OpenCV's output order of cv2.findContours(..., RETR_TREE, ...) for the borders of one image, from the
    `info` rows of plh_contour_boxes (scan position, hole flag, own key, parent key, ...): pre-order of the border
    tree, siblings in REVERSE order of discovery by the raster scan.  Returns (order, parents): indices into
    `info`, and each contour's hierarchy parent as a position in that order (-1 = top level)."""
    info = np.asarray(info)
    disc = np.argsort(info[:, 0], kind="stable")          # discovery order = scan position
    slot_of_key = {int(info[k, 2]): int(k) for k in disc}
    children = {}
    for k in disc:
        children.setdefault(slot_of_key.get(int(info[k, 3]), -1) if info[k, 3] >= 0 else -1, []).append(int(k))
    order = []
    stack = [(-1, iter(reversed(children.get(-1, []))))]
    while stack:
        try:
            k = next(stack[-1][1])
        except StopIteration:
            stack.pop()
            continue
        order.append(k)
        stack.append((k, iter(reversed(children.get(k, [])))))
    pos = {k: n for n, k in enumerate(order)}
    parents = [(-1 if info[k, 3] < 0 else pos[slot_of_key[int(info[k, 3])]]) for k in order]
    return order, parents


def contour_boxes(mask, ratio_w=1.0, ratio_h=1.0, max_contours=1024, ordered=True, return_contours=False):
    """The contour path of test.py:182-218 on the GPU: cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE)
    (hole borders included, quirk Q14) -> per contour np.int0(cv2.boxPoints(cv2.minAreaRect(.))) -> x4 ->
    /ratio_w, /ratio_h (truncating) -> order_points (the form test.py:217 writes; ordered=False: the box as
    test.py:193-201 leaves it).

    mask uint8 [H,W] or [B,H,W] (numpy or CUDA tensor; the output of pixel_detect).  Returns, per image, an
    int32 array [n,4,2] in OpenCV's contour order — and with return_contours=True also the list of contours
    (int32 [m,1,2], CHAIN_APPROX_SIMPLE points) like cv2.findContours returns them."""
    m, np_in = head.to_device(mask, torch.uint8)
    single = m.dim() == 2
    if single:
        m = m.unsqueeze(0)
    out = head.contour_boxes_raw(m, ratio_w, ratio_h, max_contours, want_points=return_contours)
    n = out["n_contours"].cpu().numpy()
    if (n > max_contours).any():
        raise ValueError("an image has %d contours; raise max_contours (=%d)" % (int(n.max()), max_contours))
    info = out["info"].cpu().numpy()
    boxes = out["boxes" if ordered else "raw_boxes"].cpu().numpy()
    pts = out["points"].cpu().numpy() if return_contours else None
    res_b, res_c = [], []
    for b in range(len(n)):
        order, _ = cv_contour_order(info[b, : n[b]])
        res_b.append(boxes[b, order].reshape(-1, 4, 2))
        if return_contours:
            res_c.append([pts[b, info[b, k, 4]: info[b, k, 4] + info[b, k, 5]].reshape(-1, 1, 2).copy() for k in order])
    if not np_in:
        res_b = [torch.as_tensor(x, device=m.device) for x in res_b]
    if single:
        return (res_b[0], res_c[0]) if return_contours else res_b[0]
    return (res_b, res_c) if return_contours else res_b


def order_points(pts):
    """test.py:24-35 — host-side ordering of one 4-point box (tl, tr, br, bl)."""
    pts = np.asarray(pts)
    x_sorted = pts[np.argsort(pts[:, 0]), :]
    left_most, right_most = x_sorted[:2, :], x_sorted[2:, :]
    left_most = left_most[np.argsort(left_most[:, 1]), :]
    tl, bl = left_most
    d = np.sqrt(((right_most.astype(np.float64) - tl.astype(np.float64)) ** 2).sum(1))
    br, tr = right_most[np.argsort(d)[::-1], :]
    return np.array([tl, tr, br, bl], dtype="int32")


def sort_poly(p):
    """test.py:37-43."""
    p = np.asarray(p)
    min_axis = np.argmin(np.sum(p, axis=1))
    p = p[[min_axis, (min_axis + 1) % 4, (min_axis + 2) % 4, (min_axis + 3) % 4]]
    if abs(p[0, 0] - p[1, 0]) > abs(p[0, 1] - p[1, 1]):
        return p
    return p[[0, 3, 2, 1]]


def write_result_txt(path, boxes):
    """ICDAR result line format of test_pixellink_fast.py:209-217 / test.py:212-218."""
    with open(path, "w") as f:
        for box in boxes:
            box = np.asarray(box)
            f.write("{},{},{},{},{},{},{},{}\r\n".format(box[0, 0], box[0, 1], box[1, 0], box[1, 1],
                                                        box[2, 0], box[2, 1], box[3, 0], box[3, 1]))
