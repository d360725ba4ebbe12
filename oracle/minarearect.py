"""Oracle: step-by-step restatement of OpenCV's minAreaRect / boxPoints on
integer point sets (un-vendored dependency of the reference).

TEST INFRASTRUCTURE — see oracle/__init__.py.

The reference calls ``cv2.minAreaRect`` → ``cv2.boxPoints`` → ``np.int0`` per
component (test_pixellink_fast.py:199-200; test.py:190-191).  OpenCV is not in
/root/reference; its published algorithm (imgproc convhull.cpp Sklansky scan +
rotcalipers.cpp rotating calipers, fp32) is restated here in scalar Python with
explicit float32 rounding so that the CUDA kernel (csrc/decode.cu) has a
line-by-line model, and it is PINNED differentially against the container's
``cv2`` 4.13.0 by tests/test_minarearect.py (random + adversarial point sets).
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def _sign(v):
    return (v > 0) - (v < 0)


def _sklansky(pts, order, start, end, nsign, sign2):
    """OpenCV convhull.cpp Sklansky_ over sorted points; returns stack of sorted positions."""
    incr = 1 if end > start else -1
    pprev, pcur, pnext = start, start + incr, start + 2 * incr
    P = lambda i: pts[order[i]]
    if start == end or (P(start)[0] == P(end)[0] and P(start)[1] == P(end)[1]):
        return [start]
    stack = [pprev, pcur, pnext]
    stacksize = 3
    end += incr
    while pnext != end:
        cury = P(pcur)[1]
        nexty = P(pnext)[1]
        by = nexty - cury
        if _sign(by) != nsign:
            ax = P(pcur)[0] - P(pprev)[0]
            bx = P(pnext)[0] - P(pcur)[0]
            ay = cury - P(pprev)[1]
            convexity = ay * bx - ax * by
            if _sign(convexity) == sign2 and (ax != 0 or ay != 0):
                pprev = pcur
                pcur = pnext
                pnext += incr
                if stacksize < len(stack):
                    stack[stacksize] = pnext
                else:
                    stack.append(pnext)
                stacksize += 1
            else:
                if pprev == start:
                    pcur = pnext
                    stack[1] = pcur
                    pnext += incr
                    stack[2] = pnext
                else:
                    stack[stacksize - 2] = pnext
                    pcur = pprev
                    pprev = stack[stacksize - 4]
                    stacksize -= 1
        else:
            pnext += incr
            stack[stacksize - 1] = pnext
    stacksize -= 1
    return stack[:stacksize]


def convex_hull_cv(points, clockwise=False, index_shift=True):
    """cv::convexHull(points, clockwise, returnPoints=true) on int points.

    ``points`` [n,2] integer (x,y).  Returns hull indices into ``points`` in
    OpenCV's output order.
    """
    pts = [(int(p[0]), int(p[1])) for p in points]
    total = len(pts)
    if total == 0:
        return []
    order = sorted(range(total), key=lambda i: (pts[i][0], pts[i][1], i))
    miny_ind = maxy_ind = 0
    for i in range(1, total):
        y = pts[order[i]][1]
        if pts[order[miny_ind]][1] > y:
            miny_ind = i
        if pts[order[maxy_ind]][1] < y:
            maxy_ind = i
    hull = []
    if pts[order[0]] == pts[order[total - 1]]:
        return [order[0]]
    tl = _sklansky(pts, order, 0, maxy_ind, -1, 1)
    tr = _sklansky(pts, order, total - 1, maxy_ind, -1, -1)
    if not clockwise:
        tl, tr = tr, tl
    for i in range(len(tl) - 1):
        hull.append(order[tl[i]])
    for i in range(len(tr) - 1, 0, -1):
        hull.append(order[tr[i]])
    stop_idx = tr[1] if len(tr) > 2 else (tl[len(tl) - 2] if len(tl) > 2 else -1)
    bl = _sklansky(pts, order, 0, miny_ind, 1, -1)
    br = _sklansky(pts, order, total - 1, miny_ind, 1, 1)
    if clockwise:
        bl, br = br, bl
    bl_count, br_count = len(bl), len(br)
    if stop_idx >= 0:
        if bl_count > 2:
            check_idx = bl[1]
        elif bl_count + br_count > 2:
            check_idx = br[2 - bl_count]
        else:
            check_idx = -1
        if check_idx == stop_idx or (check_idx >= 0 and pts[order[check_idx]] == pts[order[stop_idx]]):
            bl_count = min(bl_count, 2)
            br_count = min(br_count, 2)
    for i in range(bl_count - 1):
        hull.append(order[bl[i]])
    for i in range(br_count - 1, 0, -1):
        hull.append(order[br[i]])
    nout = len(hull)
    # cyclic shift towards a monotone index sequence
    if nout >= 3 and index_shift:
        min_idx = max_idx = 0
        lt = 0
        broke = False
        for i in range(1, nout):
            idx = hull[i]
            lt += hull[i - 1] < idx
            if lt > 1 and lt <= i - 2:
                broke = True
                break
            if idx < hull[min_idx]:
                min_idx = i
            if idx > hull[max_idx]:
                max_idx = i
        mmdist = abs(max_idx - min_idx)
        if (mmdist == 1 or mmdist == nout - 1) and (lt <= 1 or lt >= nout - 2):
            ascending = (max_idx + 1) % nout == min_idx
            i0 = min_idx if ascending else max_idx
            j = i0
            if i0 > 0:
                tmp = []
                ok = True
                for i in range(nout):
                    curr = hull[j]
                    tmp.append(curr)
                    nj = j + 1 if j + 1 < nout else 0
                    nxt = hull[nj]
                    if i < nout - 1 and (ascending != (curr < nxt)):
                        ok = False
                        break
                    j = nj
                if ok:
                    hull = tmp
    return hull


def convex_hull_giftwrap(points):
    """The same hull as ``convex_hull_cv(points, clockwise=False, index_shift=False)`` for DISTINCT
    points, computed without the sequential Sklansky scans (this is what the CUDA decode uses,
    csrc/rect.cuh): start at the lexicographic maximum (x, then y); the successor of p is the point q
    with every other point on the clockwise side of p->q in (x, y-down) coordinates, the farthest
    one among collinear candidates."""
    pts = [(int(p[0]), int(p[1])) for p in points]
    n = len(pts)
    if n == 1:
        return [0]
    start = max(range(n), key=lambda i: (pts[i][0], pts[i][1]))

    def succ(i):
        best = -1
        for j in range(n):
            if j == i:
                continue
            if best < 0:
                best = j
                continue
            ax, ay = pts[best][0] - pts[i][0], pts[best][1] - pts[i][1]
            bx, by = pts[j][0] - pts[i][0], pts[j][1] - pts[i][1]
            c = ax * by - ay * bx
            if c < 0 or (c == 0 and (ax * bx + ay * by) > 0 and bx * bx + by * by > ax * ax + ay * ay):
                best = j
        return best

    hull, cur = [start], start
    while True:
        nx = succ(cur)
        if nx == start or len(hull) > n:
            break
        hull.append(nx)
        cur = nx
    return hull


def _rotating_calipers_minarea(pts):
    """rotcalipers.cpp rotatingCalipers(..., CALIPERS_MINAREARECT) on fp32 hull points.

    Returns out[6] (corner px,py; vec1; vec2) as float32.
    """
    n = len(pts)
    px = [f32(p[0]) for p in pts]
    py = [f32(p[1]) for p in pts]
    vx = [f32(0)] * n
    vy = [f32(0)] * n
    inv = [f32(0)] * n
    left = bottom = right = top = 0
    left_x = right_x = px[0]
    top_y = bottom_y = py[0]
    pt0x, pt0y = px[0], py[0]
    for i in range(n):
        if pt0x < left_x:
            left_x, left = pt0x, i
        if pt0x > right_x:
            right_x, right = pt0x, i
        if pt0y > top_y:
            top_y, top = pt0y, i
        if pt0y < bottom_y:
            bottom_y, bottom = pt0y, i
        nx, ny = (px[i + 1], py[i + 1]) if i + 1 < n else (px[0], py[0])
        dx = float(nx) - float(pt0x)
        dy = float(ny) - float(pt0y)
        vx[i] = f32(dx)
        vy[i] = f32(dy)
        inv[i] = f32(1.0 / math.sqrt(dx * dx + dy * dy))
        pt0x, pt0y = nx, ny
    orientation = f32(0)
    ax, ay = float(vx[n - 1]), float(vy[n - 1])
    for i in range(n):
        bx, by = float(vx[i]), float(vy[i])
        convexity = ax * by - ay * bx
        if convexity != 0:
            orientation = f32(1) if convexity > 0 else f32(-1)
            break
        ax, ay = bx, by
    assert orientation != 0
    base_a, base_b = orientation, f32(0)
    seq = [bottom, right, top, left]
    minarea = f32(np.finfo(np.float32).max)
    buf = None
    for k in range(n):
        # rotated edge vectors; pick the one that comes first when rotating (exact cross products)
        rv = [
            (vx[seq[0]], vy[seq[0]]),
            (vy[seq[1]], f32(-vx[seq[1]])),    # rotate90CW
            (f32(-vx[seq[2]]), f32(-vy[seq[2]])),  # rotate180
            (f32(-vy[seq[3]]), vx[seq[3]]),    # rotate90CCW
        ]
        main = 0
        for i in range(1, 4):
            # firstVecIsRight(rv[i], rv[main]): rotate90CW(vec1) . vec2 < 0
            tx, ty = rv[i][1], f32(-rv[i][0])
            if f32(f32(tx * rv[main][0]) + f32(ty * rv[main][1])) < 0:
                main = i
        pindex = seq[main]
        lead_x = f32(vx[pindex] * inv[pindex])
        lead_y = f32(vy[pindex] * inv[pindex])
        if main == 0:
            base_a, base_b = lead_x, lead_y
        elif main == 1:
            base_a, base_b = lead_y, f32(-lead_x)
        elif main == 2:
            base_a, base_b = f32(-lead_x), f32(-lead_y)
        else:
            base_a, base_b = f32(-lead_y), lead_x
        seq[main] += 1
        if seq[main] == n:
            seq[main] = 0
        dx = f32(px[seq[1]] - px[seq[3]])
        dy = f32(py[seq[1]] - py[seq[3]])
        width = f32(f32(dx * base_a) + f32(dy * base_b))
        dx = f32(px[seq[2]] - px[seq[0]])
        dy = f32(py[seq[2]] - py[seq[0]])
        height = f32(f32(f32(-dx) * base_b) + f32(dy * base_a))
        area = f32(width * height)
        if area <= minarea:
            minarea = area
            buf = (seq[3], base_a, width, base_b, height, seq[0])
    li, A1, w, B1, h, bi = buf
    A2, B2 = f32(-B1), A1
    C1 = f32(f32(A1 * px[li]) + f32(py[li] * B1))
    C2 = f32(f32(A2 * px[bi]) + f32(py[bi] * B2))
    idet = f32(f32(1) / f32(f32(A1 * B2) - f32(A2 * B1)))
    ox = f32(f32(f32(C1 * B2) - f32(C2 * B1)) * idet)
    oy = f32(f32(f32(A1 * C2) - f32(A2 * C1)) * idet)
    return [ox, oy, f32(A1 * w), f32(B1 * w), f32(A2 * h), f32(B2 * h)]


def min_area_rect_cv(points):
    """cv::minAreaRect on integer points [n,2] -> ((cx,cy),(w,h),angle) float32.

    OpenCV 4.13 conventions measured in this container (tests/test_minarearect.py):
    hull with clockwise=false; the calipers' first side vector has an angle a in
    [0, 90] degrees; the rectangle is reported with angle in [-90, 0): a - 90 with
    width/height swapped, except a == 90 exactly which is reported as -90 unswapped.
    The angle is carried in double until the final cast.
    """
    hull_idx = convex_hull_cv(points, clockwise=False)
    hp = [(f32(points[i][0]), f32(points[i][1])) for i in hull_idx]
    n = len(hp)
    cx = cy = w = h = f32(0)
    angle = 0.0
    if n > 2:
        out = _rotating_calipers_minarea(hp)
        cx = f32(out[0] + f32(f32(out[2] + out[4]) * f32(0.5)))
        cy = f32(out[1] + f32(f32(out[3] + out[5]) * f32(0.5)))
        w = f32(math.sqrt(float(out[2]) * float(out[2]) + float(out[3]) * float(out[3])))
        h = f32(math.sqrt(float(out[4]) * float(out[4]) + float(out[5]) * float(out[5])))
        angle = math.atan2(float(out[3]), float(out[2]))
    elif n == 2:
        cx = f32(f32(hp[0][0] + hp[1][0]) * f32(0.5))
        cy = f32(f32(hp[0][1] + hp[1][1]) * f32(0.5))
        dx = float(f32(hp[1][0] - hp[0][0]))
        dy = float(f32(hp[1][1] - hp[0][1]))
        w = f32(math.sqrt(dx * dx + dy * dy))
        h = f32(0)
        angle = math.atan2(dy, dx)
    elif n == 1:
        cx, cy = hp[0]
    angle = angle * 180.0 / math.pi
    # report in [-90, 0): fold by 180 into [-90, 90), then a >= 0 -> a - 90 with w/h swapped
    if angle >= 90.0:
        angle -= 180.0
    elif angle < -90.0:
        angle += 180.0
    if angle >= 0.0:
        angle -= 90.0
        w, h = h, w
    return (cx, cy), (w, h), f32(angle)


def box_points_cv(rect):
    """cv::boxPoints / RotatedRect::points -> [4,2] float32."""
    (cx, cy), (w, h), angle = rect
    _angle = float(angle) * math.pi / 180.0
    b = f32(f32(math.cos(_angle)) * f32(0.5))
    a = f32(f32(math.sin(_angle)) * f32(0.5))
    p0x = f32(f32(cx - f32(a * h)) - f32(b * w))
    p0y = f32(f32(cy + f32(b * h)) - f32(a * w))
    p1x = f32(f32(cx + f32(a * h)) - f32(b * w))
    p1y = f32(f32(cy - f32(b * h)) - f32(a * w))
    p2x = f32(f32(f32(2) * cx) - p0x)
    p2y = f32(f32(f32(2) * cy) - p0y)
    p3x = f32(f32(f32(2) * cx) - p1x)
    p3y = f32(f32(f32(2) * cy) - p1y)
    return np.array([[p0x, p0y], [p1x, p1y], [p2x, p2y], [p3x, p3y]], np.float32)


def min_area_box_int(points):
    """np.int0(cv2.boxPoints(cv2.minAreaRect(points))) restated (truncation toward zero)."""
    return np.trunc(box_points_cv(min_area_rect_cv(points))).astype(np.int64)
