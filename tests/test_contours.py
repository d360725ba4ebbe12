"""CPU: the contour oracle (oracle/contours.py) against the container's OpenCV — findContours(RETR_TREE,
CHAIN_APPROX_SIMPLE) is the un-vendored dependency of the reference's contour path (test.py:182)."""
import numpy as np


def random_masks(n, seed=0, hmax=40, wmax=50):
    import cv2
    rng = np.random.default_rng(seed)
    for t in range(n):
        H, W = int(rng.integers(1, hmax)), int(rng.integers(1, wmax))
        m = (rng.uniform(size=(H, W)) < rng.uniform(0.15, 0.85)).astype(np.uint8)
        if t % 3 == 0:
            m = cv2.dilate(m, np.ones((3, 3), np.uint8))
        if t % 5 == 0:
            m = cv2.erode(m, np.ones((2, 2), np.uint8))
        yield m
    yield np.zeros((5, 7), np.uint8)
    yield np.ones((5, 7), np.uint8)
    m = np.ones((9, 9), np.uint8)
    m[2:7, 2:7] = 0
    m[4, 4] = 1      # ring with an island in the hole
    yield m


def test_contours_oracle_vs_cv2():
    import cv2
    from oracle import contours as C
    n = 0
    for m in random_masks(250):
        cs, h = cv2.findContours(m.copy(), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        got, parents = C.find_contours(m)
        assert len(got) == len(cs)
        for a, b in zip(got, cs):
            assert np.array_equal(a, b)
        if len(cs):
            assert [int(v) for v in h[0][:, 3]] == parents
        # the literal marked scan finds the same borders (as a set: it reports them in discovery order)
        seq = C.find_contours_sequential(m)
        assert sorted(tuple(p) for p, _, _ in seq) == sorted(tuple(map(tuple, a.reshape(-1, 2).tolist())) for a in got)
        n += len(cs)
    assert n > 2000


def test_cv_contour_order_host_helper():
    """decode.cv_contour_order rebuilds OpenCV's output order from (scan position, hole, key, parent key) rows."""
    import cv2
    from oracle import contours as C
    from tensorflow_ocr_b200.decode import cv_contour_order
    for m in random_masks(40, seed=3):
        cs, h = cv2.findContours(m.copy(), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        got, parents = C.find_contours(m)
        # build info rows the way the kernel does: any slot order, keys = arbitrary distinct ints
        rng = np.random.default_rng(len(cs))
        # discovery positions: recompute from the sequential scan
        seq = C.find_contours_sequential(m)
        key_of = {tuple(p): k for k, (p, _, _) in enumerate(seq)}
        rows = []
        for a, par in zip(got, parents):
            k = key_of[tuple(map(tuple, a.reshape(-1, 2).tolist()))]
            pk = -1 if par < 0 else key_of[tuple(map(tuple, got[par].reshape(-1, 2).tolist()))]
            rows.append([seq[k][2], seq[k][1], 1000 + k, -1 if pk < 0 else 1000 + pk, 0, len(a)])
        rows = np.asarray(rows, np.int64).reshape(-1, 6)
        perm = rng.permutation(len(rows))
        order, par2 = cv_contour_order(rows[perm])
        assert [int(perm[k]) for k in order] == list(range(len(rows)))
        assert par2 == parents
