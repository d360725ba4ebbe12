"""CPU: oracle/minarearect.py (the model of the CUDA calipers) pinned bit-exactly against
the container's OpenCV (cv2.convexHull / minAreaRect / boxPoints / np.int0)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _blob(rng):
    H, W = int(rng.integers(8, 60)), int(rng.integers(8, 80))
    img = np.zeros((H, W), np.uint8)
    box = cv2.boxPoints(((rng.uniform(0, W), rng.uniform(0, H)), (rng.uniform(3, 50), rng.uniform(2, 15)),
                         rng.uniform(-90, 90)))
    cv2.fillPoly(img, [np.round(box).astype(np.int32)], 1)
    if rng.uniform() < 0.3:
        img &= (rng.uniform(size=img.shape) > 0.3).astype(np.uint8)
    if rng.uniform() < 0.1:
        img[:] = 0
        k = int(rng.integers(1, 6))
        img[rng.integers(0, H, k), rng.integers(0, W, k)] = 1
    return np.argwhere(img > 0)


@pytest.mark.parametrize("scale", [(4.0, 3.75), (1.0, 1.0), (2.0, 1.875)])
def test_rect_emulation_matches_cv2(scale):
    from oracle import minarearect as M
    rng = np.random.default_rng(int(scale[0] * 10))
    n = 0
    for _ in range(400):
        yx = _blob(rng)
        if len(yx) == 0:
            continue
        pts = yx.copy()
        pts[:, 0] = yx[:, 1] * scale[0]     # test_pixellink_fast.py:196-197 (int64 assignment truncates)
        pts[:, 1] = yx[:, 0] * scale[1]
        for cw in (False, True):
            hull_cv = cv2.convexHull(pts.astype(np.int32), clockwise=cw, returnPoints=False).reshape(-1)
            assert np.array_equal(hull_cv, np.array(M.convex_hull_cv(pts, cw)))
        r = cv2.minAreaRect(pts.astype(np.int32))
        me = M.min_area_rect_cv(pts)
        a = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
        b = np.array([me[0][0], me[0][1], me[1][0], me[1][1], me[2]], np.float32)
        assert np.array_equal(a, b), (a, b)
        assert np.array_equal(np.intp(cv2.boxPoints(r)), M.min_area_box_int(pts))
        n += 1
    assert n > 300


def test_rect_emulation_degenerate_sets():
    from oracle import minarearect as M
    rng = np.random.default_rng(3)
    for t in range(600):
        n = int(rng.integers(1, 5))
        if t % 2:
            p0, d = rng.integers(0, 200, 2), rng.integers(-5, 6, 2)
            pts = np.array([p0 + d * k for k in range(n)])
        else:
            pts = rng.integers(0, 30, (n, 2))
        pts = np.array(sorted(pts.tolist(), key=lambda t: (t[1], t[0])), np.int64)
        r = cv2.minAreaRect(pts.astype(np.int32))
        assert np.array_equal(np.intp(cv2.boxPoints(r)), M.min_area_box_int(pts)), pts.tolist()


def test_row_extremes_are_enough():
    """The CUDA decode feeds only each row's min-x / max-x pixel to the hull (decode.cu D5):
    the result must equal cv2 on ALL the component's pixels."""
    from oracle import minarearect as M
    rng = np.random.default_rng(9)
    for _ in range(300):
        yx = _blob(rng)
        if len(yx) == 0:
            continue
        full = np.stack([(yx[:, 1] * 4.0).astype(np.int64), (yx[:, 0] * 3.75).astype(np.int64)], 1)
        cand = []
        for y in np.unique(yx[:, 0]):
            xs = yx[yx[:, 0] == y, 1]
            cand.append((int(xs.min() * 4.0), int(y * 3.75)))
            if xs.max() != xs.min():
                cand.append((int(xs.max() * 4.0), int(y * 3.75)))
        ref = np.intp(cv2.boxPoints(cv2.minAreaRect(full.astype(np.int32))))
        assert np.array_equal(M.min_area_box_int(np.array(cand, np.int64)), ref)


def test_giftwrap_hull_equals_opencv_hull_on_distinct_points():
    """The CUDA decode computes the hull of the (distinct) row extremes by parallel gift wrapping
    (csrc/rect.cuh::min_area_box_distinct): it must be the Sklansky/OpenCV vertex sequence exactly."""
    from oracle import minarearect as M
    rng = np.random.default_rng(1)
    n = 0
    for t in range(1500):
        mode = t % 3
        if mode == 0:
            yx = _blob(rng)
            if len(yx) == 0:
                continue
            sx, sy = [(4.0, 3.75), (1.0, 1.0), (2.0, 1.875)][(t // 3) % 3]
            cand = []
            for y in np.unique(yx[:, 0]):
                xs = yx[yx[:, 0] == y, 1]
                cand.append((int(xs.min() * sx), int(y * sy)))
                if xs.max() != xs.min():
                    cand.append((int(xs.max() * sx), int(y * sy)))
            pts = np.array(cand)
        elif mode == 1:
            k = int(rng.integers(1, 6))
            p0, d = rng.integers(0, 50, 2), rng.integers(-3, 4, 2)
            if (d == 0).all():
                d = np.array([1, 0])
            pts = np.array([p0 + d * i for i in range(k)])
        else:
            pts = np.unique(rng.integers(0, 12, (int(rng.integers(1, 40)), 2)), axis=0)
        pts = np.array(sorted(pts.tolist(), key=lambda t: (t[1], t[0])))
        assert M.convex_hull_giftwrap(pts) == M.convex_hull_cv(pts, False, index_shift=False)
        hull_cv = cv2.convexHull(pts.astype(np.int32), clockwise=False, returnPoints=False).reshape(-1)
        assert np.array_equal(hull_cv, np.array(M.convex_hull_cv(pts, False)))
        n += 1
    assert n > 1200
