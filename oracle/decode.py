"""Oracle: PixelLink inference decode (numpy / OpenCV restatement).

TEST INFRASTRUCTURE — see oracle/__init__.py.  Follows:

* tool/pixellink_fn.py:120-154   pixel_detect                               D1
* test_pixellink_fast.py:95-178  link graph + DFS grouping (4s maps)        D2
  test_pixellink.py:107-181      same at full resolution, min size 200
* test_pixellink_fast.py:191-202 per-group minAreaRect/boxPoints/int0       D3
* test.py:24-43,182-201          contour path, order_points, sort_poly      D4

Two component labellers are kept:
``link_components_literal`` transcribes the reference's directed DFS with seeds in
ascending pixel order (Python-2 dict order is not reproducible), and
``link_components`` is the canonical contract the CUDA kernels implement —
weakly-connected components of the same edge set, labelled by minimum pixel index.
"""
from __future__ import annotations

import numpy as np

from .pixellink_loss import softmax2

f32 = np.float32

# channel d -> (dy, dx)   (tool/pixellink_fn.py:93-108; test_pixellink_fast.py:124-146)
NEIGHBOURS = ((0, -1), (1, -1), (-1, -1), (0, 1), (1, 1), (-1, 1), (-1, 0), (1, 0))


# --------------------------------------------------------------------------- D1
def pixel_detect(score_map, geo_map, score_map_thresh=0.8, link_thresh=0.8):
    """tool/pixellink_fn.py:120-154, vectorised: 4-D input only (quirk Q9).

    score_map [1,H,W,1] probabilities; geo_map [8,1,H,W,2] link softmax outputs.
    mask = score > thr_p, then cleared wherever any link_d[...,1] < thr_l.
    """
    score_map = np.asarray(score_map)
    geo_map = np.asarray(geo_map)
    assert score_map.ndim == 4, "reference handles 4-D input only (geo_map_solo unbound otherwise)"
    score = score_map[0, :, :, 0]
    solo = geo_map[:, 0]
    res = (score > score_map_thresh).astype(np.float64)
    for i in range(8):
        res[solo[i, :, :, 1] < link_thresh] = 0
    return res.astype(np.uint8)


# --------------------------------------------------------------------------- D2
def thresholds(pixel_logits, link_logits, pixel_thresh=0.8, link_thresh=0.9):
    """test_pixellink_fast.py:53-64,96-111: softmax scores then strict `>` thresholds.

    pixel_logits [H,W,2], link_logits [H,W,16] -> P [H,W] bool, L [H,W,8] bool.
    The thresholds are Python floats compared against fp32 arrays, i.e. fp32(thr).
    """
    p1 = softmax2(pixel_logits)[..., 1]
    H, W = p1.shape
    q1 = softmax2(np.asarray(link_logits, f32).reshape(H, W, 8, 2))[..., 1]
    return p1 > f32(pixel_thresh), q1 > f32(link_thresh)


def edge_list(P, L):
    """Directed edges (v, u) of test_pixellink_fast.py:119-150: only interior pixels
    (1..W-2, 1..H-2) that are positive emit edges; to neighbour d iff L[v,d] and P[u]."""
    H, W = P.shape
    src, dst = [], []
    inner = np.zeros_like(P)
    inner[1:H - 1, 1:W - 1] = True
    for d, (dy, dx) in enumerate(NEIGHBOURS):
        ys, xs = np.nonzero(P & inner & L[:, :, d])
        uy, ux = ys + dy, xs + dx
        ok = P[uy, ux]
        src.append(ys[ok] * W + xs[ok])
        dst.append(uy[ok] * W + ux[ok])
    return np.concatenate(src), np.concatenate(dst), inner


def link_components(P, L, min_size=10):
    """Canonical contract: weakly-connected components of the reference edge set.

    Nodes = interior positive pixels + border positive pixels hit by an edge.
    Keep size > min_size (test_pixellink_fast.py:174).  Returns (labels int32 [H,W]
    with -1 background / filtered and the component's minimum linear index
    otherwise, roots ascending, sizes).
    """
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    H, W = P.shape
    N = H * W
    src, dst, inner = edge_list(P, L)
    node = (P & inner).reshape(-1).copy()
    node[dst] = True
    g = coo_matrix((np.ones(len(src), np.int8), (src, dst)), shape=(N, N))
    _, comp = connected_components(g, directed=True, connection="weak")
    idx = np.nonzero(node)[0]                      # ascending
    labels = np.full(N, -1, np.int32)
    if len(idx):
        uniq, first, inv, counts = np.unique(comp[idx], return_index=True, return_inverse=True,
                                             return_counts=True)
        mins = idx[first]                          # first occurrence == minimum pixel index
        keep_px = counts[inv] > min_size           # test_pixellink_fast.py:174
        labels[idx[keep_px]] = mins[inv][keep_px]
        kept = counts > min_size
        order = np.argsort(mins[kept])
        roots = mins[kept][order].astype(np.int32)
        root_sizes = counts[kept][order].astype(np.int32)
    else:
        roots = np.zeros(0, np.int32)
        root_sizes = np.zeros(0, np.int32)
    return labels.reshape(H, W), roots, root_sizes


def link_components_literal(P, L, min_size=10):
    """Literal transcription of test_pixellink_fast.py:113-178 (directed DFS, seeds in
    ascending key order instead of Python-2 dict order).  Pure-Python: small maps only.
    Returns group_idx [H,W] float64 with gids 1.. like the reference."""
    H, W = P.shape
    group_idx = np.zeros(H * W)
    graph = {}
    for x in range(1, W - 1):
        for y in range(1, H - 1):
            if P[y][x]:
                nb = []
                for d, (dy, dx) in enumerate(NEIGHBOURS):
                    if L[y][x][d] and P[y + dy][x + dx]:
                        nb.append((y + dy) * W + x + dx)
                graph[y * W + x] = nb
    gid = 1

    def dfs(v):
        if group_idx[v] != 0.0:
            return []
        S = [v]
        label = []
        seen = set()
        while S:
            v = S.pop()
            if v not in seen:
                seen.add(v)
                label.append(v)
                if v in graph:
                    for e in graph[v]:
                        if group_idx[e] == 0.0:
                            S.append(e)
        return label

    for i in sorted(graph.keys()):
        index_list = dfs(i)
        if len(index_list) > min_size:
            for index in index_list:
                group_idx[index] = gid
            gid += 1
    return group_idx.reshape(H, W), gid - 1


# --------------------------------------------------------------------------- D3
def component_boxes(labels, roots, scale=(4.0, 3.75)):
    """test_pixellink_fast.py:191-202 with the container's cv2: per component,
    argwhere (row-major) -> (x*sx, y*sy) assigned into an int64 array (truncation)
    -> cv2.minAreaRect -> cv2.boxPoints -> np.int0 (np.intp on numpy 2).
    Returns boxes int64 [K,4,2] in ascending label order and rects float32 [K,5]."""
    import cv2

    boxes, rects = [], []
    for r in roots:
        xy_in_poly = np.argwhere(labels == r)
        show_xy = xy_in_poly.copy()
        show_xy[:, 0] = xy_in_poly[:, 1] * scale[0]
        show_xy[:, 1] = xy_in_poly[:, 0] * scale[1]
        rectangle = cv2.minAreaRect(show_xy.astype(np.int32))
        boxes.append(np.intp(cv2.boxPoints(rectangle)))
        rects.append([rectangle[0][0], rectangle[0][1], rectangle[1][0], rectangle[1][1], rectangle[2]])
    if not boxes:
        return np.zeros((0, 4, 2), np.int64), np.zeros((0, 5), f32)
    return np.stack(boxes).astype(np.int64), np.asarray(rects, f32)


def decode_pixellink(pixel_logits, link_logits, pixel_thresh=0.8, link_thresh=0.9, min_size=10,
                     scale=(4.0, 3.75)):
    """Whole decode for one image: logits -> (labels, boxes, sizes, rects)."""
    P, L = thresholds(pixel_logits, link_logits, pixel_thresh, link_thresh)
    labels, roots, sizes = link_components(P, L, min_size)
    boxes, rects = component_boxes(labels, roots, scale)
    return labels, boxes, sizes, rects


# --------------------------------------------------------------------------- D4
def order_points(pts):
    """test.py:24-35."""
    from scipy.spatial import distance as dist

    pts = np.asarray(pts)
    x_sorted = pts[np.argsort(pts[:, 0]), :]
    left_most = x_sorted[:2, :]
    right_most = x_sorted[2:, :]
    left_most = left_most[np.argsort(left_most[:, 1]), :]
    (tl, bl) = left_most
    D = dist.cdist(tl[np.newaxis], right_most, "euclidean")[0]
    (br, tr) = right_most[np.argsort(D)[::-1], :]
    return np.array([tl, tr, br, bl], dtype="int32")


def sort_poly(p):
    """test.py:37-43."""
    p = np.asarray(p)
    min_axis = np.argmin(np.sum(p, axis=1))
    p = p[[min_axis, (min_axis + 1) % 4, (min_axis + 2) % 4, (min_axis + 3) % 4]]
    if abs(p[0, 0] - p[1, 0]) > abs(p[0, 1] - p[1, 1]):
        return p
    return p[[0, 3, 2, 1]]


def contour_boxes(mask, ratio_w=1.0, ratio_h=1.0):
    """test.py:182-201: findContours(RETR_TREE, CHAIN_APPROX_SIMPLE) incl. hole
    contours (quirk Q14), minAreaRect/boxPoints/int0, x4, /ratio (in-place int)."""
    import cv2

    contours, _ = cv2.findContours(np.ascontiguousarray(mask, np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    boxes = []
    for c in contours:
        np_contours = np.array(np.reshape(c, [-1, 2]), dtype=np.float32)
        rectangle = cv2.minAreaRect(np_contours)
        box = np.intp(cv2.boxPoints(rectangle))
        box[:, 0] = box[:, 0] * 4
        box[:, 1] = box[:, 1] * 4
        box[:, 0] = box[:, 0] / ratio_w
        box[:, 1] = box[:, 1] / ratio_h
        boxes.append(box)
    return boxes
