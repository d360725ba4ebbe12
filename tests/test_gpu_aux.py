"""Parity of the dice head, EAST loss and restore_rectangle kernels (through the C ABI)."""
import numpy as np
import pytest

from util import TOL, rel_err

pytestmark = pytest.mark.gpu


def test_dice_coefficient(golden_dir, cuda_dev):
    import torch
    from tensorflow_ocr_b200.nets import model
    g = np.load(golden_dir + "/dice_coefficient.npz")
    assert rel_err(model.dice_coefficient(g["t"], g["p"], g["m"]), g["loss"]) <= TOL
    p = torch.tensor(g["p"], device=cuda_dev, requires_grad=True)
    l = model.dice_coefficient(torch.tensor(g["t"], device=cuda_dev), p, torch.tensor(g["m"], device=cuda_dev))
    l.backward()
    assert rel_err(l.item(), g["loss"]) <= TOL
    assert rel_err(p.grad.cpu().numpy(), g["grad"]) <= TOL


def test_dice_head_golden_and_oracle(golden_dir, cuda_dev):
    import torch
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200.nets import model_vgg_16
    g = np.load(golden_dir + "/vgg16_dice_loss.npz")
    pp = torch.tensor(g["pix_prob"], device=cuda_dev, requires_grad=True)
    lp = torch.tensor(g["link_prob"], device=cuda_dev, requires_grad=True)
    l = model_vgg_16.loss(torch.tensor(g["pix_lab"], device=cuda_dev), pp, torch.tensor(g["link_lab"], device=cuda_dev),
                          lp, torch.tensor(g["training_mask"], device=cuda_dev))
    l.backward()
    assert rel_err(l.item(), g["loss"]) <= TOL
    assert rel_err(pp.grad.cpu().numpy(), g["grad_pixel"]) <= TOL
    assert rel_err(lp.grad.cpu().numpy(), g["grad_link"]) <= TOL
    # BASELINE config 5 map size, against the oracle
    rng = np.random.default_rng(1)
    B, H, W = 4, 192, 192
    t_p = (rng.uniform(size=(B, H, W, 1)) > 0.8).astype(np.float32)
    t_l = (rng.uniform(size=(B, H, W, 8)) > 0.7).astype(np.float32)
    p_p = rng.uniform(size=(B, H, W, 1)).astype(np.float32)
    p_l = rng.uniform(size=(B, H, W, 8)).astype(np.float32)
    m = (rng.uniform(size=(B, H, W, 1)) > 0.1).astype(np.float32)
    ref = O.loss_vgg16_dice(t_p, p_p, t_l, p_l, m)
    assert rel_err(model_vgg_16.loss(t_p, p_p, t_l, p_l, m), ref["loss"]) <= TOL


def test_cal_link_loss(cuda_dev):
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import synth
    from tensorflow_ocr_b200.nets import model_vgg_16
    inp = synth.make_batch(41, 2, 24, 24, "G")
    W = (inp["pix_lab"].reshape(-1) == 1).astype(np.float32)
    gt, pred = inp["link_lab"][..., 3:4], inp["link_logits"][..., 6:8]
    assert rel_err(model_vgg_16.cal_link_loss(gt, pred, W), O.cal_link_loss(gt, pred, W)) <= TOL


def test_east_loss_vs_oracle(cuda_dev):
    """E2 — parity unpinned (not in the reference; restated from upstream EAST)."""
    import torch
    from oracle import east as E
    from tensorflow_ocr_b200 import head, synth
    inp = synth.make_east_batch(4, 3, 64, 64)
    ref = E.east_loss(inp["score_gt"], inp["score_pred"], inp["geo_gt"], inp["geo_pred"], inp["training_mask"])
    t = {k: torch.as_tensor(v).to(cuda_dev) for k, v in inp.items()}
    outv, gs, gg = head.east_loss_raw(t["score_gt"], t["score_pred"], t["geo_gt"], t["geo_pred"], t["training_mask"])
    torch.cuda.synchronize()
    assert rel_err(outv[0].item(), ref["loss"]) <= TOL
    assert rel_err(gs.cpu().numpy(), ref["grad_score"]) <= TOL
    assert rel_err(gg.cpu().numpy(), ref["grad_geo"]) <= 5e-5   # cosf/sinf/logf vs numpy: a few ulp on tiny terms


def test_restore_rectangle(golden_dir, cuda_dev):
    from oracle import east as E
    from tensorflow_ocr_b200.datasets import icdar
    g = np.load(golden_dir + "/restore_rectangle.npz")
    out = icdar.restore_rectangle(g["origin"], g["geometry"])
    assert out.dtype == np.float64 and out.shape == g["out"].shape
    assert np.allclose(out, g["out"], rtol=1e-6, atol=1e-4)     # fp32 cos/sin differ by an ulp between libms
    out, idx = icdar.restore_rectangle_rbox(g["origin"], g["geometry"], return_index=True)
    a = g["geometry"][:, 4]
    expect = np.concatenate([np.nonzero(a >= 0)[0], np.nonzero(a < 0)[0]])
    assert np.array_equal(idx, expect)                           # stable partition, icdar.py:479
    assert icdar.restore_rectangle(g["origin"][:0], g["geometry"][:0]).shape == (0, 4, 2)
    # large ragged N across many scan blocks
    rng = np.random.default_rng(2)
    N = 70001
    origin = rng.uniform(0, 512, (N, 2)).astype(np.float32)
    geom = np.concatenate([rng.uniform(1, 80, (N, 4)), rng.uniform(-0.7, 0.7, (N, 1))], 1).astype(np.float32)
    ref = E.restore_rectangle_rbox(origin, geom)
    assert np.allclose(icdar.restore_rectangle(origin, geom), ref, rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("o64,g64", [(False, True), (True, False), (True, True)])
def test_restore_rectangle_float64_inputs(golden_dir, o64, g64, cuda_dev):
    """float64 inputs are computed in float64 like numpy does (not silently downcast): origins above 2^24 and a
    float64 geometry keep their digits.  Checked against the reference function itself (executed on the float64
    inputs by the oracle restatement, which numpy evaluates in the input dtype)."""
    from oracle import east as E
    from tensorflow_ocr_b200.datasets import icdar
    rng = np.random.default_rng(4)
    N = 3000
    origin = rng.uniform(0, 5e7, (N, 2)).astype(np.float64 if o64 else np.float32)
    geom = np.concatenate([rng.uniform(1, 80, (N, 4)), rng.uniform(-0.7, 0.7, (N, 1))], 1)
    geom = geom.astype(np.float64 if g64 else np.float32)
    ref = E.restore_rectangle_rbox(origin, geom)
    out = icdar.restore_rectangle(origin, geom)
    assert out.dtype == np.float64
    if g64:      # fp64 trig: device libm vs host libm, last-bit differences only
        assert np.allclose(out, ref, rtol=1e-13, atol=1e-9)
    else:        # fp32 trig differs by an ulp between libms (same bound as the fp32 test above)
        assert np.allclose(out, ref, rtol=1e-6, atol=1e-4)
    if o64:      # the origin's low digits survive: a float32 round trip of the origin would be off by > 1
        bad = icdar.restore_rectangle(origin.astype(np.float32), geom)
        assert np.abs(bad - ref).max() > 0.5 and np.abs(out - ref).max() < 1e-3


def _east_boxes(rng, n_groups, per_group):
    """Row-major-like stream of slightly jittered boxes: consecutive boxes of a group overlap."""
    polys = []
    for g in range(n_groups):
        c = rng.uniform(50, 450, 2)
        w, h = rng.uniform(30, 120), rng.uniform(10, 30)
        a = rng.uniform(-0.5, 0.5)
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        base = (np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2) @ R.T + c
        for _ in range(per_group):
            q = base + rng.normal(0, 1.0, (4, 2))
            polys.append(np.concatenate([q.reshape(-1), [rng.uniform(0.8, 1.0)]]))
    return np.asarray(polys, np.float64)


def test_lanms_vs_oracle(cuda_dev):
    """E3 — parity unpinned (not in the reference; restated from upstream EAST)."""
    from oracle import east as E
    from tensorflow_ocr_b200 import locality_aware_nms as L
    rng = np.random.default_rng(6)
    for trial in range(4):
        polys = _east_boxes(rng, int(rng.integers(1, 12)), int(rng.integers(1, 30)))
        ref = E.nms_locality(polys, 0.3)
        got = L.nms_locality(polys, 0.3)
        assert got.shape == ref.shape
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-9)
    assert L.nms_locality(np.zeros((0, 9))).shape == (0, 9)
    # batched, ragged (one empty image)
    sets = [_east_boxes(rng, 3, 10), np.zeros((0, 9)), _east_boxes(rng, 5, 4)]
    offs = np.cumsum([0] + [len(s) for s in sets]).astype(np.int32)
    res = L.nms_locality_batch(np.concatenate(sets), offs, 0.3)
    for s, r in zip(sets, res):
        ref = E.nms_locality(s, 0.3) if len(s) else np.zeros((0, 9))
        assert r.shape == ref.shape and np.allclose(r, ref, rtol=1e-9, atol=1e-9)


def test_link_labels_vs_oracle_and_golden(cuda_dev):
    """plh_link_labels (the GPU part of generate_rbox) — bit-exact vs the oracle on random id maps, and the
    drop-in generate_rbox vs the reference-executed golden."""
    import os
    import torch
    from oracle import labels as OL
    from tensorflow_ocr_b200 import head
    from tensorflow_ocr_b200.tool import pixellink_fn
    rng = np.random.default_rng(5)
    for (B, H, W) in ((1, 1, 1), (2, 3, 5), (3, 37, 61), (2, 128, 128)):
        ids = np.zeros((B, H, W), np.uint8)
        for b in range(B):
            for k in range(1, 6):
                y0, x0 = rng.integers(0, H), rng.integers(0, W)
                ids[b, y0:y0 + rng.integers(1, H + 1), x0:x0 + rng.integers(1, W + 1)] = k
        link, pix = head.link_labels_raw(torch.as_tensor(ids).cuda())
        torch.cuda.synchronize()
        for b in range(B):
            assert np.array_equal(link[b].cpu().numpy(), OL.link_labels_from_ids(ids[b]))
        assert np.array_equal(pix.cpu().numpy(), (ids != 0).astype(np.float32))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "generate_rbox.npz"))
    for ci in range(int(g["n_cases"])):
        score, link, show = pixellink_fn.generate_rbox(int(g["h%d" % ci]), int(g["w%d" % ci]), g["xs%d" % ci],
                                                       g["ys%d" % ci], g["bboxes%d" % ci], g["ignored%d" % ci])
        assert np.array_equal(score, g["score%d" % ci])
        assert np.array_equal(link, g["link%d" % ci])
        assert np.array_equal(show, g["show%d" % ci])


def test_icdar_generate_rbox(golden_dir, cuda_dev):
    """N1: datasets/icdar.py generate_rbox (the generator train.sh uses) and its [::4, ::4] subsample against the
    reference-executed golden, and the kernel against the oracle on random id maps."""
    import torch
    from oracle import labels as OL
    from tensorflow_ocr_b200 import head
    from tensorflow_ocr_b200.datasets import icdar
    g = np.load(golden_dir + "/icdar_generate_rbox.npz")
    for ci in range(int(g["n_cases"])):
        s = int(g["size%d" % ci])
        score, geo, tm = icdar.generate_rbox((s, s), g["polys%d" % ci], g["tags%d" % ci])
        assert score.dtype == np.uint8 and geo.dtype == np.float32 and tm.dtype == np.uint8
        assert np.array_equal(score, g["score%d" % ci]) and np.array_equal(geo, g["geo%d" % ci])
        assert np.array_equal(tm, g["tmask%d" % ci])
        score4, geo4, tm4 = icdar.generate_rbox_4s((s, s), g["polys%d" % ci], g["tags%d" % ci])
        assert np.array_equal(geo4, g["geo4s%d" % ci]) and np.array_equal(score4.astype(np.float32), g["score4s%d" % ci])
        assert np.array_equal(tm4.astype(np.float32), g["tmask4s%d" % ci])
    rng = np.random.default_rng(8)
    B, S = 3, 80
    first = rng.integers(0, 4, (B, S, S)).astype(np.int32)
    last = np.where(first > 0, first + rng.integers(0, 3, (B, S, S)), 0).astype(np.int32)
    link, score = head.link_labels_icdar_raw(torch.as_tensor(last).to(cuda_dev), torch.as_tensor(first).to(cuda_dev))
    for b in range(B):
        assert np.array_equal(link[b].cpu().numpy(), OL.icdar_link_labels(last[b], first[b]))
        assert np.array_equal(score[b].cpu().numpy(), (last[b] > 0).astype(np.float32))
    with pytest.raises(ValueError):
        icdar.generate_rbox((64, 96), g["polys0"], g["tags0"])


def test_east_loss_at_config4_shape(cuda_dev):
    """E2 at BASELINE config 4's shape (32 x 128x128, geometry consistent per instance like the bench's) against
    the oracle — parity unpinned (the EAST geometry loss is not in the reference)."""
    import torch
    from oracle import east as E
    from tensorflow_ocr_b200 import head, synth
    d = synth.make_east_batch(4, 32, 128, 128, consistent=True)
    ref = E.east_loss(d["score_gt"], d["score_pred"], d["geo_gt"], d["geo_pred"], d["training_mask"])
    t = {k: torch.as_tensor(v).to(cuda_dev) for k, v in d.items()}
    outv, gs, gg = head.east_loss_raw(t["score_gt"], t["score_pred"], t["geo_gt"], t["geo_pred"], t["training_mask"])
    torch.cuda.synchronize()
    assert rel_err(outv[0].item(), ref["loss"]) <= TOL
    assert rel_err(gs.cpu().numpy(), ref["grad_score"]) <= TOL
    assert rel_err(gg.cpu().numpy(), ref["grad_geo"]) <= 5e-5   # cosf/sinf/logf vs numpy: a few ulp on tiny terms


def test_rasterize_polygons_matches_reference_execution(golden_dir, cuda_dev):
    """Drawing half of the generate_rbox mirror (plh_fill_quads: cv2.fillPoly + nearest resize on the GPU) vs the
    reference-executed golden: the score map directly, the id map through the oracle's link labels."""
    import numpy as np
    from oracle import labels as OL
    from tensorflow_ocr_b200.tool import pixellink_fn
    g = np.load(golden_dir + "/generate_rbox.npz")
    for ci in range(int(g["n_cases"])):
        score, ids = pixellink_fn.rasterize_polygons(int(g["h%d" % ci]), int(g["w%d" % ci]), g["xs%d" % ci], g["ys%d" % ci])
        assert np.array_equal(score, g["score%d" % ci])
        assert ids.dtype == np.uint8 and np.array_equal(OL.link_labels_from_ids(ids), g["link%d" % ci])


@pytest.mark.parametrize("case", ["inside", "leaving", "tiny", "many"])
def test_fill_quads_vs_cv2(case, cuda_dev):
    """plh_fill_quads against the OpenCV calls it replaces: cv2.fillPoly of every polygon in order (last / first
    covering polygon, uint8 saturation, flagged polygons clearing the training mask), sampled with [::s, ::s] and
    with cv2.resize(INTER_NEAREST); polygons that leave the canvas are clipped the way cv2 clips them."""
    import cv2
    import torch
    from tensorflow_ocr_b200 import head
    rng = np.random.default_rng({"inside": 1, "leaving": 2, "tiny": 3, "many": 4}[case])
    shapes = {"inside": [(128, 128), (96, 160)], "leaving": [(100, 140), (64, 64)], "tiny": [(9, 13), (33, 21)], "many": [(150, 200)]}[case]
    for (H, W) in shapes:
        counts, quads, flags = [], [], []
        for b in range(3):
            n = {"inside": 9, "leaving": 14, "tiny": 5, "many": 300}[case] + b
            lo, hi = (0.0, 1.0) if case == "inside" else (-0.4, 1.4)
            q = []
            for k in range(n):
                if k % 4 == 3:   # arbitrary (possibly self-intersecting) quadrilateral
                    q.append(np.stack([rng.uniform(lo, hi, 4) * W, rng.uniform(lo, hi, 4) * H], -1))
                else:
                    c = np.array([rng.uniform(lo, hi) * W, rng.uniform(lo, hi) * H])
                    hw, hh, a = rng.uniform(0.02, 0.3) * W, rng.uniform(0.02, 0.15) * H, rng.uniform(-1.0, 1.0)
                    R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
                    q.append(np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T + c)
            q = np.stack(q).astype(np.float32).astype(np.int32)
            if case == "inside":
                q = np.clip(q, 0, [W - 1, H - 1]).astype(np.int32)
            counts.append(n), quads.append(q), flags.append((rng.uniform(size=n) < 0.3).astype(np.uint8))
        dq = torch.as_tensor(np.concatenate(quads)).to(cuda_dev)
        df = torch.as_tensor(np.concatenate(flags)).to(cuda_dev)
        ref = []
        for b in range(3):
            last, first, u8, tm = np.zeros((H, W), np.int32), np.zeros((H, W), np.int32), np.zeros((H, W), np.uint8), np.ones((H, W), np.uint8)
            score = np.zeros((H, W), np.float32)
            for k, quad in enumerate(quads[b]):
                cv2.fillPoly(last, quad[None], k + 1), cv2.fillPoly(u8, quad[None], k + 1), cv2.fillPoly(score, quad[None], 1.0)
                if flags[b][k]:
                    cv2.fillPoly(tm, quad[None], 0)
            for k in range(len(quads[b]) - 1, -1, -1):
                cv2.fillPoly(first, quads[b][k][None], k + 1)
            ref.append((last, first, u8, score, tm))
        want = ("last", "first", "ids_u8", "score", "training_mask")
        for stride in (1, 4):
            Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
            out = head.fill_quads_raw(dq, counts, H, W, Ho, Wo, mode=0, stride=stride, zero_flags=df, want=want)
            for b in range(3):
                for name, r in zip(want, ref[b]):
                    assert np.array_equal(out[name][b].cpu().numpy(), r[::stride, ::stride]), (case, H, W, stride, b, name)
        if H >= 4 and W >= 4:
            Ho, Wo = H // 4, W // 4
            out = head.fill_quads_raw(dq, counts, H, W, Ho, Wo, mode=1, zero_flags=df, want=want)
            for b in range(3):
                for name, r in zip(want, ref[b]):
                    rr = cv2.resize(r.astype(np.float32) if r.dtype == np.int32 else r, (Wo, Ho), interpolation=cv2.INTER_NEAREST)
                    assert np.array_equal(out[name][b].cpu().numpy(), rr.astype(r.dtype)), (case, H, W, "resize", b, name)


def test_fill_quads_argument_errors(cuda_dev):
    import torch
    from tensorflow_ocr_b200 import head
    q = torch.zeros((2, 4, 2), dtype=torch.int32, device=cuda_dev)
    with pytest.raises(ValueError):
        head.fill_quads_raw(q, [3], 32, 32, 32, 32)                       # counts do not add up
    with pytest.raises(ValueError):
        head.fill_quads_raw(q, [2], 32, 32, 32, 4096)                     # output wider than 2048
    with pytest.raises(ValueError):
        head.fill_quads_raw(q, [2], 32, 32, 32, 32, mode=0, stride=2)     # grid reaches past the canvas
    out = head.fill_quads_raw(q[:0], [0], 16, 16, 16, 16, want=("last", "training_mask"))   # no polygons: empty maps
    assert int(out["last"].abs().sum()) == 0 and int(out["training_mask"].min()) == 1
