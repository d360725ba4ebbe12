"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the ground-truth label generation of the reference,
``tool/pixellink_fn.py:9-111`` (valid_link + generate_rbox).  Pinned against the reference's own
source executed in the build container: tests/golden/generate_rbox.npz (made by
tests/golden/make_golden.py, which only patches the two Python-2 idioms `h/4` and `zip`).
"""
from __future__ import annotations

import numpy as np

# channel order of res_link_map (tool/pixellink_fn.py:90-105) as (dx, dy)
LINK_DIRS = [(-1, 0), (-1, 1), (-1, -1), (1, 0), (1, 1), (1, -1), (0, -1), (0, 1)]


def link_labels_from_ids(poly_mask: np.ndarray) -> np.ndarray:
    """tool/pixellink_fn.py:81-109: for every pixel of every polygon, valid_link() in the 8 directions.

    valid_link (:9-47): 1.0 on the map border (:10-11), else 1.0 iff the neighbour has the polygon's id."""
    m = np.asarray(poly_mask)
    h, w = m.shape
    out = np.zeros((h, w, 8), np.float32)
    inside = m != 0
    border = np.zeros((h, w), bool)
    border[0, :] = border[-1, :] = True
    border[:, 0] = border[:, -1] = True
    for d, (dx, dy) in enumerate(LINK_DIRS):
        nb = np.zeros_like(m)
        ys0, ys1 = max(0, -dy), h - max(0, dy)
        xs0, xs1 = max(0, -dx), w - max(0, dx)
        nb[ys0:ys1, xs0:xs1] = m[ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
        out[..., d] = (inside & (border | (nb == m))).astype(np.float32)
    return out


def link_labels_from_ids_loop(poly_mask: np.ndarray) -> np.ndarray:
    """Literal per-pixel form of the same loop (small maps only): the check of the vectorised one."""
    m = np.asarray(poly_mask)
    h, w = m.shape
    out = np.zeros((h, w, 8), np.float32)
    for v in np.unique(m[m != 0]):
        for y, x in np.argwhere(m == v):
            for d, (dx, dy) in enumerate(LINK_DIRS):
                if x == w - 1 or y == h - 1 or x == 0 or y == 0:
                    out[y, x, d] = 1.0
                else:
                    out[y, x, d] = 1.0 if m[y + dy, x + dx] == v else 0.0
    return out


def generate_rbox(h, w, xs, ys, bboxes, ignored):
    """tool/pixellink_fn.py:53-111 (fillPoly at full size, INTER_NEAREST to (w/4, h/4), link loop)."""
    import cv2
    assert len(xs) == len(ignored)
    h, w = int(h), int(w)
    new_h, new_w = h // 4, w // 4
    score_map = np.zeros((h, w), np.float32)
    poly_mask = np.zeros((h, w), np.uint8)
    show_bboxes = np.zeros((200, 4), np.float32)
    xs = np.asarray(xs, np.float32)
    ys = np.asarray(ys, np.float32)
    for idx in range(xs.shape[0]):
        pts = list(zip(xs[idx, :] * w, ys[idx, :] * h))
        show_bboxes[idx, :] = np.asarray(bboxes)[idx, :]
        poly = np.array([pts], np.int32)
        cv2.fillPoly(score_map, poly, 1.0)
        cv2.fillPoly(poly_mask, poly, idx + 1)
    res_score = cv2.resize(score_map, (new_w, new_h), interpolation=cv2.INTER_NEAREST)
    pm = cv2.resize(poly_mask, (new_w, new_h), interpolation=cv2.INTER_NEAREST)
    return res_score, link_labels_from_ids(pm), show_bboxes, pm


# ---------------------------------------------------------------------------------------------------------------
# datasets/icdar.py:83-105 valid_link + :486-539 generate_rbox — the EAST-fork generator (quirk Q17), pinned by
# tests/golden/icdar_generate_rbox.npz (the reference's own source executed).
ICDAR_DIRS = [(0, -1), (1, -1), (-1, -1), (0, 1), (1, 1), (-1, 1), (-1, 0), (1, 0)]   # (dx, dy) AS THE CODE MOVES


def icdar_link_labels(last_ids, first_ids):
    """Link labels from the last-cover / first-cover polygon index maps (see csrc/aux.cu for why these two maps
    state the order-dependent loop of generate_rbox exactly).  Square maps."""
    L = np.asarray(last_ids)
    F = np.asarray(first_ids)
    h, w = L.shape
    assert h == w
    out = np.zeros((h, w, 8), np.float32)
    ys, xs = np.nonzero(L)
    k = L[ys, xs]
    edge = (xs == h - 1) | (ys == w - 1)
    for d, (dx, dy) in enumerate(ICDAR_DIRS):
        qx, qy = (xs + dx) % w, (ys + dy) % h          # -1 wraps like numpy; the far side never occurs (edge rule)
        f = F[qy, qx]
        out[ys, xs, d] = np.where(edge, 1.0, ((f != 0) & (f <= k)).astype(np.float32))
    return out


def icdar_generate_rbox(im_size, polys, tags, min_text_size=10):
    """datasets/icdar.py:486-539 restated with the two-map form."""
    import cv2
    h, w = im_size
    last = np.zeros((h, w), np.int32)
    first = np.zeros((h, w), np.int32)
    training_mask = np.ones((h, w), np.uint8)
    polys = np.asarray(polys)
    quads = [np.asarray(p).astype(np.int32)[np.newaxis] for p in polys]
    for k, (poly, tag) in enumerate(zip(polys, tags)):
        cv2.fillPoly(last, quads[k], k + 1)
        poly_h = min(np.linalg.norm(poly[0] - poly[3]), np.linalg.norm(poly[1] - poly[2]))
        poly_w = min(np.linalg.norm(poly[0] - poly[1]), np.linalg.norm(poly[2] - poly[3]))
        if min(poly_h, poly_w) < min_text_size or tag:
            cv2.fillPoly(training_mask, quads[k], 0)
    for k in range(len(quads) - 1, -1, -1):
        cv2.fillPoly(first, quads[k], k + 1)
    return (last > 0).astype(np.uint8), icdar_link_labels(last, first), training_mask
