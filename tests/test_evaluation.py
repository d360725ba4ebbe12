"""N4 evaluation code on the CPU: the oracle against the reference-executed golden (tests/golden/evaluation.npz,
make_golden.py::golden_evaluation), the interval statement of cv2's filled contour against cv2 itself, and the
host-side metrics mirror."""
import os

import numpy as np
import pytest

from oracle import evaluation as oe

GOLD = os.path.join(os.path.dirname(__file__), "golden", "evaluation.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_jaccard_and_matching_equal_the_reference(gold):
    for ci in range(int(gold["n_cases"])):
        bboxes, gxs, gys, gi = gold["bboxes%d" % ci], gold["gxs%d" % ci], gold["gys%d" % ci], gold["gignored%d" % ci]
        jac = np.stack([oe.np_bboxes_jaccard(b, gxs, gys) for b in bboxes])
        assert jac.dtype == np.float32 and np.array_equal(jac, gold["jaccard%d" % ci])
        n_g, tp, fp = oe.bboxes_matching(bboxes, gxs, gys, gi, 0.5)
        assert n_g == int(gold["n_gbboxes%d" % ci])
        assert np.array_equal(tp, gold["tp%d" % ci]) and np.array_equal(fp, gold["fp%d" % ci])
        pre, rec = oe.precision_recall(n_g, tp, fp)
        assert pre == gold["precision%d" % ci] and rec == gold["recall%d" % ci]
        np.testing.assert_array_equal(oe.fmean(pre, rec), gold["fmean%d" % ci])


def test_interval_statement_equals_the_reference_jaccard(gold):
    """quad_jaccard_rows (no cv2: what the CUDA kernel computes) is bit-identical to the reference's mask path."""
    for ci in range(int(gold["n_cases"])):
        bboxes, gxs, gys = gold["bboxes%d" % ci], gold["gxs%d" % ci], gold["gys%d" % ci]
        if bboxes[:, 0::2].max() > 400:   # the big canvases are covered on the GPU; keep the CPU suite short
            bboxes = bboxes[:3]
        jac = np.stack([oe.quad_jaccard_rows(b, gxs, gys) for b in bboxes])
        assert np.array_equal(jac, gold["jaccard%d" % ci][:len(bboxes)])


def test_filled_polygon_rows_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for t in range(3000):
        kind = t % 5
        if kind == 0:
            pts = rng.integers(0, 60, (4, 2))
        elif kind == 1:
            pts = rng.integers(0, 12, (4, 2))          # many degenerate shapes: repeated points, segments
        elif kind == 2:
            c, w, h, a = rng.uniform(20, 200, 2), rng.uniform(1, 80), rng.uniform(1, 30), rng.uniform(-3.2, 3.2)
            R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
            pts = np.maximum((np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2 @ R.T + c), 0)
        elif kind == 3:
            pts = rng.integers(0, 300, (4, 2))         # self-intersecting quadrilaterals
        else:
            pts = rng.integers(0, 40, (int(rng.integers(3, 9)), 2))
        pts = np.asarray(pts).astype(np.int32)
        shape = (int(pts[:, 1].max()) + 10, int(pts[:, 0].max()) + 10)
        ref = np.zeros(shape, np.uint8)
        cv2.drawContours(ref, [pts.reshape(-1, 1, 2)], -1, 1, -1)
        got = oe.mask_from_rows(oe.filled_quad_rows([tuple(int(v) for v in p) for p in pts]), shape)
        assert np.array_equal(ref, got), pts.tolist()


def test_clipped_lines_and_polygons_vs_cv2():
    """Vertices outside the image: cv::clipLine, and the edge table built from the clipped segments."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(6)
    for t in range(3000):
        p = rng.integers(-40, 90, (2, 2))
        w, h = 60, 50
        ref = np.zeros((h, w), np.uint8)
        cv2.line(ref, tuple(int(v) for v in p[0]), tuple(int(v) for v in p[1]), 1, 1, 8)
        vis, q0, q1 = oe.clip_line(w, h, tuple(int(v) for v in p[0]), tuple(int(v) for v in p[1]))
        got = oe.mask_from_rows({y: [r] for y, r in oe.line_row_runs(q0, q1).items()} if vis else {}, (h, w))
        assert np.array_equal(ref, got), p.tolist()
    for t in range(3000):
        pts = (rng.integers(-30, 80, (4, 2)) if t % 2 else rng.integers(-8, 40, (4, 2))).astype(np.int32)
        if t % 3 == 0:   # the reference's mask: 10 more than the largest coordinate, so only x < 0 / y < 0 clip
            h, w = max(int(pts[:, 1].max()), 1) + 10, max(int(pts[:, 0].max()), 1) + 10
        else:
            h, w = 50, 60
        ref = np.zeros((h, w), np.uint8)
        cv2.drawContours(ref, [pts.reshape(-1, 1, 2)], -1, 1, -1)
        got = oe.mask_from_rows(oe.filled_quad_rows([tuple(int(v) for v in p) for p in pts], w, h), (h, w))
        assert np.array_equal(ref, got), (pts.tolist(), h, w)


def test_interval_jaccard_with_negative_coordinates():
    rng = np.random.default_rng(8)
    for t in range(40):
        bbox = rng.integers(-15, 60, 8)
        gxs, gys = rng.integers(-15, 70, (3, 4)), rng.integers(-15, 50, (3, 4))
        if max(bbox[0::2].max(), gxs.max()) < 0 or max(bbox[1::2].max(), gys.max()) < 0:
            continue
        assert np.array_equal(oe.quad_jaccard_rows(bbox, gxs, gys), oe.np_bboxes_jaccard(bbox, gxs, gys)), (bbox, gxs, gys)


def test_metrics_mirror_equals_the_reference(gold):
    from tensorflow_ocr_b200.tool import metrics
    state = None
    tot_n, tot_tp, tot_fp = 0, [], []
    for ci in range(int(gold["n_cases"])):
        n_g, tp, fp = gold["n_gbboxes%d" % ci], gold["tp%d" % ci], gold["fp%d" % ci]
        pre, rec = metrics.precision_recall(n_g, tp, fp)
        assert pre.dtype == np.float32 and pre == gold["precision%d" % ci] and rec == gold["recall%d" % ci]
        np.testing.assert_array_equal(metrics.fmean(pre, rec), gold["fmean%d" % ci])
        (v_n, v_tp, v_fp), state = metrics.streaming_tp_fp_arrays(n_g, tp, fp, state=state)
        tot_n += int(n_g)
        tot_tp.append(tp), tot_fp.append(fp)
        assert v_n == tot_n and np.array_equal(v_tp, np.concatenate(tot_tp)) and np.array_equal(v_fp, np.concatenate(tot_fp))
    # the streaming arrays feed precision_recall at the end of an evaluation (tool/metrics.py:31-80)
    pre, rec = metrics.precision_recall(*state.value())
    o_pre, o_rec = oe.precision_recall(tot_n, np.concatenate(tot_tp), np.concatenate(tot_fp))
    assert pre == o_pre and rec == o_rec
    assert metrics.precision_recall(0, np.zeros(0, bool), np.zeros(0, bool)) == (0.0, 0.0)   # safe_divide


def test_filled_polygon_properties_hypothesis():
    """Property form of the two tests above: any quadrilateral with vertices in or around any small canvas."""
    cv2 = pytest.importorskip("cv2")
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    @hyp.settings(max_examples=400, deadline=None)
    @hyp.given(st.lists(st.tuples(st.integers(-40, 110), st.integers(-40, 110)), min_size=4, max_size=4),
               st.integers(1, 90), st.integers(1, 90))
    def check(pts, w, h):
        arr = np.array(pts, np.int32)
        ref = np.zeros((h, w), np.uint8)
        cv2.drawContours(ref, [arr.reshape(-1, 1, 2)], -1, 1, -1)
        got = oe.mask_from_rows(oe.filled_quad_rows([tuple(p) for p in pts], w, h), (h, w))
        assert np.array_equal(ref, got)
        vis, q0, q1 = oe.clip_line(w, h, pts[0], pts[1])
        line = np.zeros((h, w), np.uint8)
        cv2.line(line, pts[0], pts[1], 1, 1, 8)
        got = oe.mask_from_rows({y: [r] for y, r in oe.line_row_runs(q0, q1).items()} if vis else {}, (h, w))
        assert np.array_equal(line, got)

    check()
