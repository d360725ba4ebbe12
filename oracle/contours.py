"""Oracle: cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE) restated (numpy / pure Python; small maps).

TEST INFRASTRUCTURE — see oracle/__init__.py.  The contour path of the reference (test.py:182-218) calls
OpenCV for this; OpenCV is an un-vendored dependency, so the restatement below (Suzuki & Abe border following
as OpenCV's contours.cpp implements it: scan order, start rules, search directions, CHAIN_APPROX_SIMPLE point
emission) is pinned against the container's cv2 4.13.0 in tests/test_contours.py, point sequence by point
sequence and in cv2's output order.

Two statements of the same thing are kept:
  * ``find_contours_sequential`` — the literal raster scan with border marking (the -NBD / NBD marks decide
    which later "1 -> 0" and "0 -> 1" transitions start a border);
  * ``find_contours`` — the form the CUDA kernels use: the marks are not needed once one knows that an outer
    border starts exactly at the raster-first pixel of every 8-connected foreground component and a hole border
    exactly at the left neighbour of the raster-first pixel of every 4-connected background region that does not
    touch the frame; every border is then traced independently on the unmarked binary image.
"""
from __future__ import annotations

import numpy as np

DX = (1, 1, 0, -1, -1, -1, 0, 1)      # direction codes of OpenCV (CV_INIT_3X3_DELTAS): 0 = right, counter-clockwise
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def trace_border(img, x0, y0, is_hole, mark=None):
    """icvFetchContour on a zero-padded binary image `img` (nonzero = foreground), start pixel (x0, y0) in
    padded coordinates.  Returns the CHAIN_APPROX_SIMPLE points in unpadded (x, y).  `mark`: optional int array
    receiving Suzuki's marks (2 = visited, -2 = visited with an examined 0-pixel on its right)."""
    pts = []
    s_end = s = 0 if is_hole else 4
    while True:                                    # first neighbour, searching clockwise
        s = (s - 1) & 7
        x1, y1 = x0 + DX[s], y0 + DY[s]
        if img[y1, x1] != 0 or s == s_end:
            break
    if img[y1, x1] == 0:                           # isolated pixel
        if mark is not None:
            mark[y0, x0] = -2
        return [(x0 - 1, y0 - 1)]
    x3, y3 = x0, y0
    prev_s = s ^ 4
    px, py = x0, y0
    while True:
        s_end = s
        while True:                                # next neighbour, searching counter-clockwise
            s += 1
            x4, y4 = x3 + DX[s & 7], y3 + DY[s & 7]
            if img[y4, x4] != 0:
                break
        s &= 7
        if mark is not None:
            if ((s - 1) & 0xFFFFFFFF) < s_end:     # the pixel on the right was examined and is 0
                mark[y3, x3] = -2
            elif mark[y3, x3] == 1:
                mark[y3, x3] = 2
        if s != prev_s:                            # CHAIN_APPROX_SIMPLE: a point where the direction changes
            pts.append((px - 1, py - 1))
            prev_s = s
        px += DX[s]
        py += DY[s]
        if (x4, y4) == (x0, y0) and (x3, y3) == (x1, y1):
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return pts


def find_contours_sequential(mask):
    """Literal scan (cvFindNextContour).  Returns [(points, is_hole, scan_pos)] in discovery order."""
    mask = np.asarray(mask)
    H, W = mask.shape
    img = np.zeros((H + 2, W + 2), np.int32)
    img[1:-1, 1:-1] = (mask != 0)
    out = []
    for y in range(1, H + 1):
        prev = 0
        for x in range(1, W + 1):
            p = img[y, x]
            if p != prev:
                is_hole = 0
                start = True
                if not (prev == 0 and p == 1):
                    if p != 0 or prev < 1:
                        start = False
                    else:
                        is_hole = 1
                if start:
                    out.append((trace_border(img, x - is_hole, y, is_hole, mark=img), is_hole, (y - 1) * W + x - 1))
                    p = img[y, x]
                prev = p
    return out


def _components(fg, conn8):
    """Label map (minimum linear index per component, -1 elsewhere) of a boolean map."""
    from scipy import ndimage
    st = np.ones((3, 3), int) if conn8 else np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    lab, n = ndimage.label(fg, structure=st)
    H, W = fg.shape
    idx = np.arange(H * W).reshape(H, W)
    out = np.full((H, W), -1, np.int64)
    if n:
        mins = ndimage.minimum(idx, lab, np.arange(1, n + 1)).astype(np.int64)
        out[lab > 0] = mins[lab[lab > 0] - 1]
    return out


def find_contours(mask):
    """cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0] as a list of int32 arrays [n,1,2],
    in OpenCV's order (pre-order of the border tree, siblings in reverse discovery order), plus the hierarchy
    parents.  Starts from component labels; every border traced independently."""
    mask = np.asarray(mask) != 0
    H, W = mask.shape
    img = np.zeros((H + 2, W + 2), np.int32)
    img[1:-1, 1:-1] = mask
    fg = _components(mask, True)
    bgp = _components(np.pad(~mask, 1, constant_values=True), False)      # frame ring joins everything that touches the border
    outside = bgp[0, 0]
    bg = bgp[1:-1, 1:-1]
    items = []   # (scan position, is_hole, key, parent key, points)
    for y in range(H):
        for x in range(W):
            i = y * W + x
            if mask[y, x] and fg[y, x] == i:                   # raster-first pixel of a foreground component
                left = bg[y, x - 1] if x > 0 else outside
                parent = None if left == outside else ("h", int(left))
                items.append((i, 0, ("o", i), parent, trace_border(img, x + 1, y + 1, 0)))
            elif not mask[y, x] and bg[y, x] != outside:
                yy, xx = divmod(int(bg[y, x]) - (W + 2) - 1, W + 2)        # region root in unpadded coordinates
                if (yy, xx) == (y, x):                         # raster-first pixel of a hole: border starts at its left neighbour
                    items.append((i, 1, ("h", int(bg[y, x])), ("o", int(fg[y, x - 1])), trace_border(img, x, y + 1, 1)))
    items.sort(key=lambda t: t[0])
    keys = {t[2]: k for k, t in enumerate(items)}
    parent = [(-1 if t[3] is None else keys[t[3]]) for t in items]
    children = {}
    for k, p in enumerate(parent):
        children.setdefault(p, []).append(k)
    order = []

    def visit(p):
        for k in reversed(children.get(p, [])):
            order.append(k)
            visit(k)
    visit(-1)
    pos = {k: n for n, k in enumerate(order)}
    contours = [np.asarray(items[k][4], np.int32).reshape(-1, 1, 2) for k in order]
    parents = [(-1 if parent[k] < 0 else pos[parent[k]]) for k in order]
    return contours, parents
