"""Parity of the CUDA decode (through the C ABI) against the oracle / OpenCV: bit-exact
labels (canonical min-index) and integer box corners."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _decode(inp, **kw):
    """Runs BOTH forms of the decode (tiled multi-kernel = the default / resident: one 8-CTA cluster per image
    with the map as bit planes in shared memory, whenever a strip fits) and asserts that they agree bit for
    bit before handing the default's result to the caller's oracle check."""
    import torch
    from tensorflow_ocr_b200 import head
    dev = torch.device("cuda", 0)
    pl, ll = torch.as_tensor(inp["pix_logits"]).to(dev), torch.as_tensor(inp["link_logits"]).to(dev)
    res = []
    for form in ("tiled", "resident"):
        try:
            out = head.decode_raw(pl, ll, head.DecodeConfig(form=form, **kw))
        except ValueError:           # PLH_E_SHAPE: a strip of this map does not fit in shared memory
            assert form == "resident" and inp["pix_logits"].shape[1] * inp["pix_logits"].shape[2] > 300 * 300
            res.append(res[0])
            continue
        torch.cuda.synchronize()
        res.append({k: v.cpu().numpy() for k, v in out.items()})
    a, b = res
    assert np.array_equal(a["labels"], b["labels"]) and np.array_equal(a["n_boxes"], b["n_boxes"])
    K = a["boxes"].shape[1]
    for i, n in enumerate(a["n_boxes"]):
        if n <= K:   # beyond the capacity the written subset is unspecified
            for k in ("boxes", "comp", "rects"):
                assert np.array_equal(a[k][i, :n], b[k][i, :n]), (k, i)
    return a


def _check(inp, out, min_size=10, scale=(4.0, 3.75), pixel_thresh=0.8, link_thresh=0.9):
    from oracle import decode as D
    B = inp["pix_logits"].shape[0]
    nb = 0
    for b in range(B):
        lab, boxes, sizes, rects = D.decode_pixellink(inp["pix_logits"][b], inp["link_logits"][b], pixel_thresh,
                                                      link_thresh, min_size, scale)
        assert np.array_equal(out["labels"][b], lab), "labels differ in image %d" % b
        n = len(boxes)
        assert out["n_boxes"][b] == n
        assert np.array_equal(out["comp"][b, :n, 1], sizes)
        assert np.array_equal(out["comp"][b, :n, 0], np.unique(lab[lab >= 0]))
        assert np.array_equal(out["boxes"][b, :n], boxes), "box corners differ in image %d" % b
        assert np.array_equal(out["rects"][b, :n], rects), "rects differ in image %d" % b
        nb += n
    return nb


@pytest.mark.parametrize("family", ["G", "S"])
@pytest.mark.parametrize("B,H,W", [(4, 128, 128), (3, 48, 80), (2, 192, 320), (5, 33, 47)])
def test_decode_vs_oracle(family, B, H, W, cuda_dev):
    from tensorflow_ocr_b200 import synth
    inp = synth.make_batch(11, B, H, W, family, edge_images=(B >= 4))
    out = _decode(inp, max_boxes=256)
    nb = _check(inp, out)
    assert nb > 0 or H < 40


def test_decode_icdar_shaped_2s_map(cuda_dev):
    """BASELINE config 3a map size: 768x1280 network input at 2s -> 384x640 maps, box scale (2.0, 1.875)."""
    from tensorflow_ocr_b200 import synth
    inp = synth.make_batch(15, 2, 384, 640, "G")
    out = _decode(inp, scale=(2.0, 1.875), max_boxes=512)
    assert _check(inp, out, scale=(2.0, 1.875)) > 0


def test_decode_pixellink_api(cuda_dev):
    """The named drop-in: numpy in -> (labels, boxes, counts)."""
    from oracle import decode as D
    from tensorflow_ocr_b200 import synth
    from tensorflow_ocr_b200.decode import decode_pixellink
    inp = synth.make_batch(16, 3, 64, 80, "G")
    labels, boxes, counts = decode_pixellink(inp["pix_logits"], inp["link_logits"])
    for b in range(3):
        lab, bx, sizes, _ = D.decode_pixellink(inp["pix_logits"][b], inp["link_logits"][b])
        assert np.array_equal(labels[b], lab) and np.array_equal(boxes[b], bx) and np.array_equal(counts[b], sizes)
    l1, b1, c1 = decode_pixellink(inp["pix_logits"][0], inp["link_logits"][0])
    assert np.array_equal(l1, labels[0]) and np.array_equal(b1, boxes[0])


def test_decode_variants(cuda_dev):
    """full-res script constants (test_pixellink.py:177,209-215): min_size 200, no scaling;
    and a 2s configuration (scale 2.0, 1.875)."""
    from tensorflow_ocr_b200 import synth
    inp = synth.make_batch(12, 2, 160, 256, "G")
    out = _decode(inp, min_size=200, scale=(1.0, 1.0), max_boxes=64, pixel_thresh=0.6, link_thresh=0.6)
    _check(inp, out, min_size=200, scale=(1.0, 1.0), pixel_thresh=0.6, link_thresh=0.6)
    out = _decode(inp, min_size=0, scale=(2.0, 1.875), max_boxes=4096, pixel_thresh=0.7, link_thresh=0.7)
    _check(inp, out, min_size=0, scale=(2.0, 1.875), pixel_thresh=0.7, link_thresh=0.7)


@pytest.mark.parametrize("tag", ["fast", "full"])
def test_decode_link_graph_golden(golden_dir, tag, cuda_dev):
    """D2/D3 against the reference scripts' own graph + DFS + minAreaRect lines, executed by line range
    (tests/golden/make_golden.py:golden_link_graph): 192x320 / > 10 px / scale (4, 3.75) of
    test_pixellink_fast.py:111-202 and 720x1280 / > 200 px / unscaled of test_pixellink.py:113-217."""
    from util import link_graph_case
    c = link_graph_case(golden_dir, tag)
    inp = {"pix_logits": c["pix_logits"][None], "link_logits": c["link_logits"][None]}
    out = _decode(inp, min_size=c["min_size"], scale=c["scale"], max_boxes=64)
    n = c["n_groups"]
    assert out["n_boxes"][0] == n
    assert np.array_equal(out["labels"][0], c["labels"])
    assert np.array_equal(out["boxes"][0, :n], c["boxes"])


def test_decode_literal_dfs_crosscheck():
    """Oracle-only: on symmetric links without positive border pixels the canonical
    components equal the reference's literal DFS grouping."""
    from oracle import decode as D
    from tensorflow_ocr_b200 import synth
    for i in range(3):
        im = synth.make_image(13, i, 40, 56, "S")
        P, L = D.thresholds(im["pix_logits"], im["link_logits"])
        P[0, :] = P[-1, :] = False
        P[:, 0] = P[:, -1] = False
        lab, roots, sizes = D.link_components(P, L, 10)
        g, n = D.link_components_literal(P, L, 10)
        assert n == len(roots)
        for gid in range(1, n + 1):
            px = np.argwhere(g == gid)
            r = lab[px[0][0], px[0][1]]
            assert r >= 0 and np.array_equal(lab == r, g == gid)


def test_decode_asymmetric_links_vs_literal_dfs(cuda_dev):
    """Quirk Q10, quantified: on ASYMMETRIC link maps the reference's directed, seed-order dependent DFS
    (test_pixellink_fast.py:151-178, here with ascending seeds) and the canonical weakly-connected components
    this library returns are different groupings.  What always holds, and is asserted: every literal group
    lies inside ONE canonical component (a directed reachability set is weakly connected), so the canonical
    grouping never splits what the reference joins; the reverse merges are counted and reported."""
    from oracle import decode as D
    from tensorflow_ocr_b200 import synth
    n_lit = n_can = 0
    for i in range(3):
        im = synth.make_image(31, i, 48, 64, "G")          # independent link logits per direction: asymmetric
        P, L = D.thresholds(im["pix_logits"], im["link_logits"], 0.6, 0.6)
        out = _decode({"pix_logits": im["pix_logits"][None], "link_logits": im["link_logits"][None]},
                      pixel_thresh=0.6, link_thresh=0.6, min_size=0, max_boxes=2048)
        lab = out["labels"][0]
        g, n = D.link_components_literal(P, L, 0)
        for gid in range(1, n + 1):
            inside = np.unique(lab[g == gid])
            assert len(inside) == 1 and inside[0] >= 0, "a literal group straddles two canonical components"
        n_lit += n
        n_can += len(np.unique(lab[lab >= 0]))
    assert n_can <= n_lit
    print("asymmetric links: %d literal DFS groups -> %d canonical components" % (n_lit, n_can))


def test_min_area_boxes_point_limit(cuda_dev):
    """More than 2048 points in one explicit set: sentinel box instead of a shared-memory overrun."""
    import torch
    from tensorflow_ocr_b200 import head
    rng = np.random.default_rng(9)
    a = rng.integers(0, 500, (3000, 2)).astype(np.int32)
    b = rng.integers(0, 500, (100, 2)).astype(np.int32)
    b = np.array(sorted(b.tolist(), key=lambda t: (t[1], t[0])), np.int32)
    offs = np.array([0, 3000, 3100], np.int32)
    boxes, rects = head.min_area_boxes_raw(torch.as_tensor(np.concatenate([a, b])).to(cuda_dev),
                                           torch.as_tensor(offs).to(cuda_dev))
    boxes = boxes.cpu().numpy()
    import cv2
    assert (boxes[0] == np.iinfo(np.int32).min).all()
    assert np.array_equal(boxes[1], np.intp(cv2.boxPoints(cv2.minAreaRect(b))))


def test_min_area_boxes_vs_cv2(cuda_dev):
    """CUDA rotating calipers vs cv2.minAreaRect/boxPoints/np.int0 on explicit point lists."""
    import cv2
    import torch
    from tensorflow_ocr_b200 import head
    rng = np.random.default_rng(5)
    sets = []
    for t in range(400):
        mode = t % 5
        if mode == 0:
            n = int(rng.integers(1, 300))
            pts = rng.integers(0, int(rng.integers(5, 1300)), (n, 2))
        elif mode == 1:
            H, W = int(rng.integers(10, 120)), int(rng.integers(10, 160))
            img = np.zeros((H, W), np.uint8)
            cv2.ellipse(img, (W // 2, H // 2), (int(rng.integers(2, W // 2)), int(rng.integers(2, H // 2))),
                        float(rng.uniform(0, 180)), 0, 360, 1, -1)
            yx = np.argwhere(img > 0)
            pts = np.stack([yx[:, 1] * 2.0, yx[:, 0] * 1.875], 1).astype(np.int64)
        elif mode == 2:
            n = int(rng.integers(1, 5))
            p0 = rng.integers(0, 200, 2)
            d = rng.integers(-5, 6, 2)
            pts = np.array([p0 + d * k for k in range(n)])
        elif mode == 3:
            H, W = int(rng.integers(3, 60)), int(rng.integers(3, 60))
            img = np.zeros((H, W), np.uint8)
            y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W))
            img[y0:y0 + int(rng.integers(1, 10)), x0:x0 + int(rng.integers(1, 30))] = 1
            yx = np.argwhere(img > 0)
            pts = np.stack([yx[:, 1] * 4.0, yx[:, 0] * 3.75], 1).astype(np.int64)
        else:
            H, W = int(rng.integers(8, 60)), int(rng.integers(8, 80))
            img = np.zeros((H, W), np.uint8)
            box = cv2.boxPoints(((rng.uniform(0, W), rng.uniform(0, H)), (rng.uniform(3, 50), rng.uniform(2, 15)),
                                 rng.uniform(-90, 90)))
            cv2.fillPoly(img, [np.round(box).astype(np.int32)], 1)
            yx = np.argwhere(img > 0)
            pts = np.stack([yx[:, 1] * 4.0, yx[:, 0] * 3.75], 1).astype(np.int64)
        if len(pts) == 0 or len(pts) > 2048:
            continue
        pts = np.array(sorted(pts.tolist(), key=lambda t: (t[1], t[0])), np.int32).reshape(-1, 2)
        sets.append(pts)
    offs = np.zeros(len(sets) + 1, np.int32)
    offs[1:] = np.cumsum([len(s) for s in sets])
    allp = np.concatenate(sets).astype(np.int32)
    boxes, rects = head.min_area_boxes_raw(torch.as_tensor(allp).to(cuda_dev), torch.as_tensor(offs).to(cuda_dev))
    boxes, rects = boxes.cpu().numpy(), rects.cpu().numpy()
    bad = 0
    for i, s in enumerate(sets):
        r = cv2.minAreaRect(s)
        ref_box = np.intp(cv2.boxPoints(r))
        ref_rect = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
        if not (np.array_equal(boxes[i], ref_box) and np.array_equal(rects[i], ref_rect)):
            bad += 1
    assert bad == 0, "%d / %d point sets differ from cv2" % (bad, len(sets))


def test_pixel_detect(golden_dir, cuda_dev):
    from tensorflow_ocr_b200.tool import pixellink_fn
    g = np.load(golden_dir + "/pixel_detect.npz")
    assert np.array_equal(pixellink_fn.pixel_detect(g["score"], g["geo"]), g["res"])
    assert np.array_equal(pixellink_fn.pixel_detect(g["score"], g["geo"], 0.75, 0.7), g["res_075_07"])


def test_pixel_detect_test_py_layout(golden_dir, cuda_dev):
    """`decode.pixel_detect` keeps the name and input layout of test.py:45-74 (score [1,H,W,1], geo [1,H,W,16],
    channel 2i+1 = link_i score) with the semantics of tool/pixellink_fn.py:120-154 (quirk Q8: the twin in
    test.py clears two cells instead of every failing pixel); pinned on the reference-executed golden of the
    latter, re-laid-out."""
    import torch
    from tensorflow_ocr_b200.decode import pixel_detect
    g = np.load(golden_dir + "/pixel_detect.npz")
    geo = g["geo"]                                                  # [8,1,H,W,2]
    H, W = geo.shape[2], geo.shape[3]
    geo16 = np.ascontiguousarray(np.transpose(geo[:, 0], (1, 2, 0, 3)).reshape(1, H, W, 16))
    assert np.array_equal(pixel_detect(g["score"], geo16), g["res"])
    assert np.array_equal(pixel_detect(g["score"], geo16, 0.75, 0.7), g["res_075_07"])
    t = pixel_detect(torch.as_tensor(g["score"]).to(cuda_dev), torch.as_tensor(geo16).to(cuda_dev))
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), g["res"])
    with pytest.raises(ValueError):
        pixel_detect(g["score"], geo16[..., :8])


def test_fused_loss_decode_flags(cuda_dev):
    """The fused path (flags emitted by the loss kernel) decodes identically to the standalone decode."""
    import torch
    from tensorflow_ocr_b200 import head, synth
    inp = synth.make_batch(14, 6, 64, 96, "G", edge_images=True)
    t = {k: torch.as_tensor(inp[k]).to(cuda_dev) for k in ("pix_logits", "link_logits", "pix_lab", "link_lab")}
    dcfg = head.DecodeConfig(max_boxes=128)
    alone = head.decode_raw(t["pix_logits"], t["link_logits"], dcfg)
    torch.cuda.synchronize()
    n = alone["n_boxes"].cpu().numpy()
    from oracle import pixellink_loss as O
    ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
    for parallel in (False, True):   # one stream + flags from the loss kernel / two concurrent streams
        fused = head.loss_and_decode_raw(t["pix_logits"], t["link_logits"], t["pix_lab"], t["link_lab"],
                                         head.LossConfig(), dcfg, want_rects=True, parallel=parallel)
        torch.cuda.synchronize()
        for k in ("labels", "n_boxes"):
            assert torch.equal(fused[k], alone[k])
        for b in range(6):
            assert torch.equal(fused["boxes"][b, :n[b]], alone["boxes"][b, :n[b]])
        out = {k: v.cpu().numpy() for k, v in fused.items()}
        _check(inp, out)
        assert abs(out["stats"][0] - ref["loss"]) <= 1e-5 * abs(ref["loss"])


def test_decode_box_capacity_overflow(cuda_dev):
    """More components than max_boxes: n_boxes reports the true count, the label map is complete, and
    the K written boxes are K distinct components of the image (an unspecified subset)."""
    from oracle import decode as D
    from tensorflow_ocr_b200 import synth
    inp = synth.make_batch(17, 1, 128, 128, "G")
    lab, boxes, sizes, _ = D.decode_pixellink(inp["pix_logits"][0], inp["link_logits"][0], min_size=2)
    assert len(boxes) > 3
    out = _decode(inp, min_size=2, max_boxes=3)
    assert out["n_boxes"][0] == len(boxes)
    assert np.array_equal(out["labels"][0], lab)
    roots = np.unique(lab[lab >= 0])
    got_roots = out["comp"][0, :3, 0]
    assert len(set(got_roots.tolist())) == 3 and set(got_roots.tolist()) <= set(roots.tolist())
    for k in range(3):
        j = int(np.nonzero(roots == got_roots[k])[0][0])
        assert np.array_equal(out["boxes"][0, k], boxes[j])


@pytest.mark.parametrize("seed", range(6))
def test_decode_random_shapes(seed, cuda_dev):
    """Ragged map sizes (not multiples of the 32x16 tile), random thresholds and scales."""
    from tensorflow_ocr_b200 import synth
    rng = np.random.default_rng(100 + seed)
    B, H, W = int(rng.integers(1, 4)), int(rng.integers(3, 150)), int(rng.integers(3, 200))
    fam = "GS"[seed % 2]
    inp = synth.make_batch(18 + seed, B, H, W, fam)
    tp, tl = float(rng.choice([0.5, 0.7, 0.8])), float(rng.choice([0.5, 0.8, 0.9]))
    sc = [(4.0, 3.75), (1.0, 1.0), (2.0, 1.875)][seed % 3]
    ms = int(rng.integers(0, 12))
    out = _decode(inp, pixel_thresh=tp, link_thresh=tl, scale=sc, min_size=ms, max_boxes=1024)
    _check(inp, out, min_size=ms, scale=sc, pixel_thresh=tp, link_thresh=tl)


def test_decode_phase_bits_compose(cuda_dev):
    """reserved[0] bit 2 (tile pass only) followed by bit 3 (the rest) on one workspace == the whole decode."""
    import dataclasses
    import torch
    from tensorflow_ocr_b200 import _lib, head, synth
    inp = synth.make_batch(23, 3, 70, 100, "G")
    dev = torch.device("cuda", 0)
    pl, ll = torch.as_tensor(inp["pix_logits"]).to(dev), torch.as_tensor(inp["link_logits"]).to(dev)
    cfg = head.DecodeConfig(min_size=3, max_boxes=256)
    whole = {k: v.clone() for k, v in head.decode_raw(pl, ll, cfg, None, True).items()}
    ws = torch.empty(_lib.load().plh_workspace_bytes(_lib.OP_DECODE, 3, 70, 100, 256), dtype=torch.uint8, device=dev)
    out = {}
    head.decode_raw(pl, ll, dataclasses.replace(cfg, phase=4), out, True, ws)
    head.decode_raw(pl, ll, dataclasses.replace(cfg, phase=8), out, True, ws)
    torch.cuda.synchronize()
    nb = whole["n_boxes"].cpu().numpy()
    assert np.array_equal(out["n_boxes"].cpu().numpy(), nb)
    assert np.array_equal(out["labels"].cpu().numpy(), whole["labels"].cpu().numpy())
    for b in range(3):
        assert np.array_equal(out["boxes"][b, :nb[b]].cpu().numpy(), whole["boxes"][b, :nb[b]].cpu().numpy())


def test_contour_path_vs_cv2(cuda_dev):
    """N2 / D4: the contour path of test.py:182-218 on the GPU against OpenCV itself: the CHAIN_APPROX_SIMPLE point
    sequence of every border (outer and hole) in cv2's output order, and the boxes after minAreaRect / boxPoints
    / int0 / x4 / truncating division by the ratios, before and after order_points."""
    import cv2
    from oracle import decode as D
    from tensorflow_ocr_b200.decode import contour_boxes
    from test_contours import random_masks
    nc = 0
    for t, m in enumerate(random_masks(60, seed=7, hmax=70, wmax=90)):
        rw, rh = [(1.0, 1.0), (0.8, 0.75), (1.6, 1.28)][t % 3]
        cs, _ = cv2.findContours(m.copy(), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        raw, conts = contour_boxes(m, rw, rh, max_contours=4096, ordered=False, return_contours=True)
        assert len(conts) == len(cs)
        for a, b in zip(conts, cs):
            assert np.array_equal(a, b)
        ref = D.contour_boxes(m, rw, rh)                    # test.py:182-201 with the container's cv2
        assert len(ref) == len(raw)
        for a, b in zip(raw, ref):
            assert np.array_equal(a, b), (t, a.tolist(), b.tolist())
        ordered = contour_boxes(m, rw, rh, max_contours=4096)
        for a, b in zip(ordered, ref):
            assert np.array_equal(a, D.order_points(b))     # test.py:217
        nc += len(cs)
    assert nc > 1000


def test_contour_path_batched_text_masks(cuda_dev):
    """Batched, on masks shaped like the path's real input: pixel_detect outputs of synthetic PixelLink maps."""
    import cv2
    import torch
    from oracle import decode as D
    from oracle.pixellink_loss import softmax2
    from tensorflow_ocr_b200 import synth
    from tensorflow_ocr_b200.decode import contour_boxes
    B, H, W = 4, 128, 128
    inp = synth.make_batch(33, B, H, W, "G")
    masks = []
    for b in range(B):
        score = softmax2(inp["pix_logits"][b])[None, :, :, 1:2]
        link = softmax2(inp["link_logits"][b].reshape(H, W, 8, 2)).transpose(2, 0, 1, 3)[:, None]
        masks.append(D.pixel_detect(score, link, 0.6, 0.3))
    masks = np.stack(masks)
    got = contour_boxes(torch.as_tensor(masks).to(cuda_dev), 0.9, 0.8, max_contours=4096)
    n = 0
    for b in range(B):
        ref = D.contour_boxes(masks[b], 0.9, 0.8)
        assert len(ref) == len(got[b])
        for a, r in zip(got[b].cpu().numpy(), ref):
            assert np.array_equal(a, D.order_points(r))
        n += len(ref)
    assert n > 20
