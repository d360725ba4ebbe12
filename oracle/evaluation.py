"""CPU oracle for the evaluation code — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package; the product
(tensorflow_ocr_b200/) never does.

Restates
  * tool/bboxes.py:252-282  np_bboxes_jaccard  — Jaccard of one detected quadrilateral against G ground-truth
    quadrilaterals by RASTERISING both onto a 0/1 mask with cv2.drawContours(thickness = -1) and counting pixels;
  * tool/bboxes.py:158-246  bboxes_matching    — greedy Pascal-VOC matching of detections (in the given order)
    against ground truth with an `ignored` flag;
  * tool/metrics.py:31-85   streaming_tp_fp_arrays / precision_recall / fmean.

`util.img.{points_to_contours, black, draw_contours}` come from the un-vendored `util` package (dengdan/pylib);
they are one-line wrappers of numpy / cv2 (`np.zeros(shape, np.uint8)`, `cv2.drawContours(img, contours, idx,
color, border_width)`), restated as such.  cv2 itself is the reference's dependency (container: 4.13.0).

Pinned (tests/test_oracle_golden.py) against tests/golden/evaluation.npz, produced by executing the reference's
own function bodies (tests/golden/make_golden.py::golden_evaluation).

`filled_quad_rows` is the interval statement of cv2's filled-contour rasterisation that the CUDA kernel
(csrc/evalbox.cu) follows; tests/test_evaluation.py pins it against cv2.drawContours on random polygons.
"""
from __future__ import annotations

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT


# ----------------------------------------------------------------------------- tool/bboxes.py:252-282
def np_bboxes_jaccard(bbox, gxs, gys):
    """bbox (8,) x0,y0..x3,y3; gxs, gys (G,4).  float32 (G,).  tool/bboxes.py:252-282 with util.img spelled out."""
    import cv2

    bbox_points = np.reshape(bbox, (4, 2))
    cnts = [np.asarray(bbox_points, np.int32).reshape(-1, 1, 2)]            # util.img.points_to_contours
    xmax = max(np.max(bbox_points[:, 0]), np.max(gxs)) + 10                 # :258-261
    ymax = max(np.max(bbox_points[:, 1]), np.max(gys)) + 10
    mask = np.zeros((int(ymax), int(xmax)), np.uint8)                        # util.img.black
    bbox_mask = mask.copy()
    cv2.drawContours(bbox_mask, cnts, -1, 1, -1)                             # :265
    jaccard = np.zeros((len(gxs),), dtype=np.float32)
    for gt_idx, gt_bbox in enumerate(zip(gxs, gys)):                         # :268-281
        gt_mask = mask.copy()
        gt_bbox = np.transpose(gt_bbox)
        cv2.drawContours(gt_mask, [np.asarray(gt_bbox, np.int32).reshape(-1, 1, 2)], -1, 1, -1)
        intersect = np.sum(bbox_mask * gt_mask)
        union = np.sum(bbox_mask + gt_mask >= 1)
        jaccard[gt_idx] = intersect * 1.0 / union
    return jaccard


# ----------------------------------------------------------------------------- tool/bboxes.py:158-246
def bboxes_matching(bboxes, gxs, gys, gignored, matching_threshold=0.5, jaccard_fn=np_bboxes_jaccard):
    """bboxes (N,8) in score order; gxs, gys (G,4); gignored (G,).  -> n_gbboxes, tp (N,) bool, fp (N,) bool."""
    gignored = np.asarray(gignored).astype(bool)
    n_gbboxes = int(np.count_nonzero(~gignored))                             # :180
    gmatch = np.zeros(gignored.shape, bool)                                  # :182
    n = len(bboxes)
    tp = np.zeros(n, bool)
    fp = np.zeros(n, bool)
    for i in range(n):                                                       # :198-226 (parallel_iterations = 1)
        jaccard = jaccard_fn(bboxes[i], gxs, gys)
        idxmax = int(np.argmax(jaccard))                                     # first maximum
        match = jaccard[idxmax] > matching_threshold                         # :208 strict
        existing = gmatch[idxmax]
        not_ignored = not gignored[idxmax]
        tp[i] = not_ignored and match and not existing                       # :215
        fp[i] = not_ignored and (existing or not match)                      # :218
        if not_ignored and match:                                            # :222-223
            gmatch[idxmax] = True
    return n_gbboxes, tp, fp


# ----------------------------------------------------------------------------- tool/metrics.py
def precision_recall(num_gbboxes, tp, fp):
    """tool/metrics.py:66-80: fp32 sums, safe_divide (0 where the denominator is <= 0, tool/math.py:27-41)."""
    tp = np.float32(np.sum(np.asarray(tp).astype(np.float32)))
    fp = np.float32(np.sum(np.asarray(fp).astype(np.float32)))
    n = np.float32(num_gbboxes)
    recall = np.float32(tp / n) if n > 0 else np.float32(0)
    precision = np.float32(tp / np.float32(tp + fp)) if (tp + fp) > 0 else np.float32(0)
    return precision, recall


def fmean(pre, rec):
    """tool/metrics.py:82-85 (no guard: 0/0 -> nan, as the reference)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float32(2) * np.float32(pre) * np.float32(rec) / (np.float32(pre) + np.float32(rec))


class StreamingTpFp:
    """tool/metrics.py:31-63: the three local variables and their update op."""

    def __init__(self):
        self.num_gbboxes = np.int32(0)
        self.tp = np.zeros((0,), bool)
        self.fp = np.zeros((0,), bool)

    def update(self, num_gbboxes, tp, fp):
        self.num_gbboxes = np.int32(self.num_gbboxes + np.sum(np.asarray(num_gbboxes).astype(np.int32)))
        self.tp = np.concatenate([self.tp, np.asarray(tp).astype(bool).reshape(-1)])
        self.fp = np.concatenate([self.fp, np.asarray(fp).astype(bool).reshape(-1)])
        return self.num_gbboxes, self.tp, self.fp


# ----------------------------------------------------------------------------- cv2's filled contour as row intervals
def _tdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def line_row_runs(p0, p1):
    """Pixels of cv::line(8-connected, thickness 1) between integer points as {row: (xmin, xmax)}.

    LineIterator(leftToRight = true): the walk starts at the endpoint with the smaller x; Bresenham with
    err0 = dx - 2 dy on the major axis, so the minor coordinate after i major steps is
    m_i = (2 * dminor * i + dmajor - 1) // (2 * dmajor)."""
    (x0, y0), (x1, y1) = p0, p1
    dx, dy = x1 - x0, y1 - y0
    if dx < 0:
        x0, y0, dx, dy = x1, y1, -dx, -dy
    sy = 1
    if dy < 0:
        dy, sy = -dy, -1
    runs = {}
    if dy > dx:
        for i in range(dy + 1):
            m = (2 * dx * i + dy - 1) // (2 * dy)
            runs[y0 + sy * i] = (x0 + m, x0 + m)
    else:
        for i in range(dx + 1):
            m = (2 * dy * i + dx - 1) // (2 * dx) if dx else 0
            y = y0 + sy * m
            a, b = runs.get(y, (x0 + i, x0 + i))
            runs[y] = (min(a, x0 + i), max(b, x0 + i))
    return runs


def clip_line(w, h, p1, p2):
    """cv::clipLine(Size(w, h), pt1, pt2) (imgproc/drawing.cpp), as restated from the OpenCV sources and pinned
    against cv2.line in tests/test_evaluation.py: the y side first, then x, each intersection computed in double
    and truncated.  -> (visible, pt1', pt2')."""
    (x1, y1), (x2, y2) = p1, p2
    right, bottom = w - 1, h - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1, c1 = a, 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2, c2 = a, 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def filled_quad_rows(pts, w=None, h=None):
    """cv2.drawContours(thickness=-1) / fillPoly of one polygon with integer vertices on a w x h image, as
    {row: [(x1, x2), ...]} inclusive intervals (possibly overlapping).  w, h None: vertices are non-negative and the
    image holds them all (no clipping).

    = the outline drawn with cv::line (clipped against the image first), plus, for every row y in [ymin, ymax)
    of the non-horizontal edges, the spans between consecutive pairs of the sorted edge abscissae
    x_e(y) = x_e(y_top) + (y - y_top) * dx_e (16.16 fixed point, dx_e by truncating division), from
    ceil(x_left) to floor(x_right), cut to the image.  An edge with an endpoint outside the image takes its
    abscissae (and, unless the clipped segment is horizontal, its ordinates) from the clipped segment for the slope
    and is extrapolated back to its own top row."""
    n = len(pts)
    if w is None:
        w = max(p[0] for p in pts) + 1
        h = max(p[1] for p in pts) + 1
    rows, edges = {}, []
    for i in range(n):
        p0, p1 = pts[i - 1], pts[i]
        vis, q0, q1 = clip_line(w, h, p0, p1)
        if vis:
            for y, r in line_row_runs(q0, q1).items():
                rows.setdefault(y, []).append(r)
        if p0[1] == p1[1]:
            continue
        c0x, c0y, c1x, c1y = p0[0] << XY_SHIFT, p0[1], p1[0] << XY_SHIFT, p1[1]
        if not (0 <= p0[0] < w and 0 <= p0[1] < h and 0 <= p1[0] < w and 0 <= p1[1] < h):
            if q0[1] != q1[1]:
                c0y, c1y = q0[1], q1[1]
            c0x, c1x = q0[0] << XY_SHIFT, q1[0] << XY_SHIFT
        dxe = _tdiv(c1x - c0x, c1y - c0y)
        edges.append((p0[1], p1[1], c0x + (p0[1] - c0y) * dxe, dxe) if p0[1] < p1[1] else
                     (p1[1], p0[1], c1x + (p1[1] - c1y) * dxe, dxe))
    if edges:
        for y in range(max(min(e[0] for e in edges), 0), min(max(e[1] for e in edges), h)):
            xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
            for k in range(0, len(xs) - 1, 2):
                x1, x2 = max((xs[k] + XY_ONE - 1) >> XY_SHIFT, 0), min(xs[k + 1] >> XY_SHIFT, w - 1)
                if x1 <= x2:
                    rows.setdefault(y, []).append((x1, x2))
    return rows


def mask_from_rows(rows, shape):
    m = np.zeros(shape, np.uint8)
    for y, iv in rows.items():
        for a, b in iv:
            m[y, a:b + 1] = 1
    return m


def quad_jaccard_rows(bbox, gxs, gys):
    """np_bboxes_jaccard through filled_quad_rows (no cv2): what the CUDA kernel computes.  The mask of
    tool/bboxes.py:258-262 is (ymax + 10, xmax + 10): only the x < 0 / y < 0 sides ever clip."""
    bp = np.reshape(bbox, (4, 2))
    w = int(max(np.max(bp[:, 0]), np.max(gxs)) + 10)
    h = int(max(np.max(bp[:, 1]), np.max(gys)) + 10)
    ra = filled_quad_rows([tuple(int(v) for v in p) for p in bp], w, h)
    ma = mask_from_rows(ra, (h, w))
    out = np.zeros((len(gxs),), np.float32)
    for g in range(len(gxs)):
        rb = filled_quad_rows([(int(gxs[g][k]), int(gys[g][k])) for k in range(4)], w, h)
        mb = mask_from_rows(rb, (h, w))
        out[g] = np.sum(ma & mb) * 1.0 / np.sum(ma | mb)
    return out
