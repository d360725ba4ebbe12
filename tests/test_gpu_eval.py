"""N4 on the GPU: tool.bboxes mirrors (csrc/evalbox.cu through the C ABI) against the reference-executed golden
and against the oracle on larger random sets."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "evaluation.npz")


def _random_image(rng, size, G, D):
    def quad(c, w, h, a):
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        return np.maximum((np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2 @ R.T + c), 0).astype(np.int32)
    gts = np.stack([quad(rng.uniform(0.1, 0.9, 2) * size, rng.uniform(0.03, 0.3) * size[0], rng.uniform(0.02, 0.12) * size[1],
                         rng.uniform(-0.6, 0.6)) for _ in range(G)])
    dets = []
    for d in range(D):
        if rng.uniform() < 0.7:
            g = int(rng.integers(0, G))
            dets.append(np.maximum(gts[g] + rng.normal(0, 0.006 * size[0] * (1 + 5 * (rng.uniform() < 0.3)), (4, 2)), 0).astype(np.int32))
        else:
            dets.append(quad(rng.uniform(0.1, 0.9, 2) * size, rng.uniform(0.03, 0.2) * size[0], rng.uniform(0.02, 0.1) * size[1],
                             rng.uniform(-1.5, 1.5)))
    return np.stack(dets).reshape(-1, 8), gts[:, :, 0].copy(), gts[:, :, 1].copy(), (rng.uniform(size=G) < 0.25).astype(np.int32)


def test_jaccard_and_matching_equal_the_reference_golden():
    from tensorflow_ocr_b200.tool import bboxes as tb
    gold = np.load(GOLD)
    for ci in range(int(gold["n_cases"])):
        bboxes, gxs, gys, gi = gold["bboxes%d" % ci], gold["gxs%d" % ci], gold["gys%d" % ci], gold["gignored%d" % ci]
        for i in (0, len(bboxes) - 1):
            j = tb.np_bboxes_jaccard(bboxes[i], gxs, gys)
            assert j.dtype == np.float32 and np.array_equal(j, gold["jaccard%d" % ci][i])
            assert np.array_equal(tb.bboxes_jaccard(bboxes[i], gxs, gys), j)
        n_g, tp, fp = tb.bboxes_matching(bboxes, gxs, gys, gi, matching_threshold=0.5)
        assert int(n_g) == int(gold["n_gbboxes%d" % ci])
        assert tp.dtype == bool and np.array_equal(tp, gold["tp%d" % ci]) and np.array_equal(fp, gold["fp%d" % ci])


def test_full_jaccard_matrix_equals_the_golden():
    import torch
    from tensorflow_ocr_b200 import head
    gold = np.load(GOLD)
    n = int(gold["n_cases"])
    dets = [gold["bboxes%d" % ci].reshape(-1, 4, 2) for ci in range(n)]
    gts = [np.stack([gold["gxs%d" % ci], gold["gys%d" % ci]], -1) for ci in range(n)]
    ign = [gold["gignored%d" % ci].astype(np.uint8) for ci in range(n)]
    dev = torch.device("cuda", 0)
    out = head.bboxes_matching_raw(torch.as_tensor(np.concatenate(dets)).to(dev), torch.as_tensor(np.concatenate(gts)).to(dev),
                                   [len(d) for d in dets], [len(g) for g in gts], torch.as_tensor(np.concatenate(ign)).to(dev))
    jac = out["jaccard"].cpu().numpy()
    for ci in range(n):
        lo, hi = out["pair_off"][ci], out["pair_off"][ci + 1]
        assert np.array_equal(jac[lo:hi].reshape(len(dets[ci]), len(gts[ci])), gold["jaccard%d" % ci]), ci
        d0, d1 = out["det_off"][ci], out["det_off"][ci + 1]
        assert np.array_equal(out["tp"].cpu().numpy()[d0:d1].astype(bool), gold["tp%d" % ci])
        assert np.array_equal(out["fp"].cpu().numpy()[d0:d1].astype(bool), gold["fp%d" % ci])
    assert np.array_equal(out["n_gbboxes"].cpu().numpy(), [int(gold["n_gbboxes%d" % ci]) for ci in range(n)])


@pytest.mark.parametrize("size,G,D", [((160, 90), 12, 30), ((1280, 720), 25, 60), ((320, 192), 1, 5)])
def test_batch_vs_oracle(size, G, D):
    from oracle import evaluation as oe
    from tensorflow_ocr_b200.tool import bboxes as tb
    rng = np.random.default_rng(G * 1000 + D)
    imgs = [_random_image(rng, size, G if b else max(G // 2, 1), D + b) for b in range(3)]
    n, tps, fps = tb.bboxes_matching_batch([im[0] for im in imgs], [im[1] for im in imgs], [im[2] for im in imgs],
                                           [im[3] for im in imgs], matching_threshold=0.5)
    for b, (bboxes, gxs, gys, gi) in enumerate(imgs):
        o_n, o_tp, o_fp = oe.bboxes_matching(bboxes, gxs, gys, gi, 0.5)
        assert int(n[b]) == o_n and np.array_equal(tps[b], o_tp) and np.array_equal(fps[b], o_fp), b
        assert G == 1 or o_tp.sum() + o_fp.sum() > 0
    for i in range(0, len(imgs[1][0]), 7):   # spot-check the Jaccard rows against the cv2 mask path
        assert np.array_equal(tb.np_bboxes_jaccard(imgs[1][0][i], imgs[1][1], imgs[1][2]),
                              oe.np_bboxes_jaccard(imgs[1][0][i], imgs[1][1], imgs[1][2]))


def test_degenerate_and_identical_quads():
    from oracle import evaluation as oe
    from tensorflow_ocr_b200.tool import bboxes as tb
    gxs = np.array([[10, 50, 50, 10], [7, 7, 7, 7], [5, 40, 40, 5], [10, 60, 60, 10], [0, 0, 3, 3]], np.int32)
    gys = np.array([[10, 10, 30, 30], [7, 7, 7, 7], [5, 9, 9, 5], [10, 40, 10, 40], [0, 2, 2, 0]], np.int32)
    for bbox in ([10, 10, 50, 10, 50, 30, 10, 30], [7, 7, 7, 7, 7, 7, 7, 7], [5, 5, 40, 9, 40, 9, 5, 5], [10, 10, 60, 40, 60, 10, 10, 40],
                 [0, 0, 200, 0, 200, 100, 0, 100]):
        j = tb.np_bboxes_jaccard(np.array(bbox, np.int32), gxs, gys)
        assert np.array_equal(j, oe.np_bboxes_jaccard(np.array(bbox, np.int32), gxs, gys)), bbox
    assert tb.np_bboxes_jaccard(np.array([10, 10, 50, 10, 50, 30, 10, 30]), gxs, gys)[0] == 1.0


def test_negative_coordinates_are_clipped_like_cv2():
    """Boxes sticking out of the image at the top / left (minAreaRect corners near the border do)."""
    from oracle import evaluation as oe
    from tensorflow_ocr_b200.tool import bboxes as tb
    rng = np.random.default_rng(17)
    n = 0
    for t in range(60):
        lo = -25 if t % 2 else -6
        bboxes = rng.integers(lo, 90, (4, 8)).astype(np.int32)
        gxs, gys = rng.integers(lo, 100, (5, 4)).astype(np.int32), rng.integers(lo, 70, (5, 4)).astype(np.int32)
        gi = (rng.uniform(size=5) < 0.2).astype(np.int32)
        for b in bboxes:
            assert np.array_equal(tb.np_bboxes_jaccard(b, gxs, gys), oe.np_bboxes_jaccard(b, gxs, gys)), (b, gxs, gys)
            n += 1
        o_n, o_tp, o_fp = oe.bboxes_matching(bboxes, gxs, gys, gi, 0.3)
        g_n, g_tp, g_fp = tb.bboxes_matching(bboxes, gxs, gys, gi, matching_threshold=0.3)
        assert int(g_n) == o_n and np.array_equal(g_tp, o_tp) and np.array_equal(g_fp, o_fp)
    assert n == 240


def test_no_detections():
    """An image without detections: empty TP / FP arrays, the ground-truth count still reported."""
    from tensorflow_ocr_b200.tool import bboxes as tb
    gxs = np.array([[1, 5, 5, 1], [10, 20, 20, 10]]); gys = np.array([[1, 1, 5, 5], [10, 10, 14, 14]])
    n, tp, fp = tb.bboxes_matching(np.zeros((0, 8), np.int32), gxs, gys, np.array([0, 1]))
    assert int(n) == 1 and tp.shape == (0,) and fp.shape == (0,)
    n, tps, fps = tb.bboxes_matching_batch([np.zeros((0, 8), np.int32), np.array([[1, 1, 5, 1, 5, 5, 1, 5]])], [gxs, gxs], [gys, gys],
                                           [np.array([0, 0]), np.array([0, 0])])
    assert list(n) == [2, 2] and tps[0].shape == (0,) and list(tps[1]) == [True] and list(fps[1]) == [False]


def test_argument_errors():
    from tensorflow_ocr_b200.tool import bboxes as tb
    gxs = np.array([[1, 5, 5, 1]]); gys = np.array([[1, 1, 5, 5]])
    with pytest.raises(ValueError):
        tb.np_bboxes_jaccard(np.array([1 << 20, 0, 5, 0, 5, 5, 0, 5]), gxs, gys)     # coordinate out of range
    with pytest.raises(ValueError):
        tb.np_bboxes_jaccard(np.arange(8), np.zeros((0, 4), np.int32), np.zeros((0, 4), np.int32))   # no ground truth
    with pytest.raises(ValueError):
        tb.bboxes_matching(np.arange(16).reshape(2, 8), gxs, gys, np.array([0, 1]))   # gignored length
