"""CPU: the oracle against the golden vectors produced by executing the reference's own
source (tests/golden/make_golden.py) — this is what pins the oracle."""
import numpy as np

from util import TOL, rel_err


def test_example_softmax_known_answer(golden_dir):
    """example.py:5,12 — the only known answer in the reference: softmax([a, a+1])."""
    from oracle import pixellink_loss as O
    g = np.load(golden_dir + "/example_softmax.npz")
    y = O.softmax2(g["x"])
    assert np.allclose(y, g["y"], rtol=1e-6)
    assert np.allclose(y[..., 0], 0.26894142, atol=1e-7) and np.allclose(y[..., 1], 0.73105858, atol=1e-7)


def test_model_loss_golden(golden_dir):
    from oracle import pixellink_loss as O
    g = np.load(golden_dir + "/model_loss_b14.npz")
    r = O.loss_model(g["pix_lab"], g["pix_logits"], g["link_lab"], g["link_logits"])
    assert rel_err(r["loss"], g["loss"]) <= TOL
    assert np.array_equal(r["ohem_mask"], g["ohem_mask"])
    assert rel_err(r["grad_pixel"], g["grad_pixel"]) <= TOL
    assert rel_err(r["grad_link"], g["grad_link"]) <= TOL


def test_model_loss_golden_nan(golden_dir):
    from oracle import pixellink_loss as O
    g = np.load(golden_dir + "/model_loss_b14_nopos.npz")
    r = O.loss_model(g["pix_lab"], g["pix_logits"], g["link_lab"], g["link_logits"])
    assert np.isnan(r["loss"]) and np.isnan(g["loss"])
    assert rel_err(r["grad_pixel"], g["grad_pixel"]) == 0.0
    assert np.isnan(r["grad_link"]).all() and np.isnan(g["grad_link"]).all()


def test_vgg16_goldens(golden_dir):
    from oracle import pixellink_loss as O
    g = np.load(golden_dir + "/vgg16_ohem_loss.npz")
    r = O.ohem_loss_vgg16(g["pix_lab"], g["pix_logits"], g["link_lab"], g["link_logits"])
    assert rel_err(r["loss"], g["loss"]) <= TOL
    assert rel_err(r["grad_pixel"], g["grad_pixel"]) <= TOL
    assert rel_err(r["grad_link"], g["grad_link"]) <= TOL
    g = np.load(golden_dir + "/vgg16_dice_loss.npz")
    r = O.loss_vgg16_dice(g["pix_lab"], g["pix_prob"], g["link_lab"], g["link_prob"], g["training_mask"])
    assert rel_err(r["loss"], g["loss"]) <= TOL
    assert rel_err(r["grad_pixel"], g["grad_pixel"]) <= TOL
    assert rel_err(r["grad_link"], g["grad_link"]) <= TOL
    g = np.load(golden_dir + "/dice_coefficient.npz")
    l, gr, _, _ = O.dice_coefficient(g["t"], g["p"], g["m"], True)
    assert rel_err(l, g["loss"]) <= TOL and rel_err(gr, g["grad"]) <= TOL


def test_pixellink_build_loss_golden(golden_dir):
    from oracle import pixellink_loss as O
    g = np.load(golden_dir + "/pixellink_build_loss.npz")
    r = O.build_loss_pixellink(g["pix_logits"], g["link_logits"], g["pix_lab"], g["link_lab"])
    assert rel_err(r["losses"][0], g["losses"][0]) <= TOL
    assert rel_err(r["losses"][1], g["losses"][1]) <= TOL
    assert np.array_equal(r["ohem_mask"], g["ohem_mask"])
    assert rel_err(r["grad_pixel"], g["grad_pixel"]) <= TOL
    assert rel_err(r["grad_link"], g["grad_link"]) <= TOL


def test_pixel_detect_golden(golden_dir):
    from oracle import decode as D
    g = np.load(golden_dir + "/pixel_detect.npz")
    assert np.array_equal(D.pixel_detect(g["score"], g["geo"]), g["res"])
    assert np.array_equal(D.pixel_detect(g["score"], g["geo"], 0.75, 0.7), g["res_075_07"])


def test_restore_rectangle_golden(golden_dir):
    from oracle import east as E
    g = np.load(golden_dir + "/restore_rectangle.npz")
    out = E.restore_rectangle_rbox(g["origin"], g["geometry"])
    assert out.dtype == np.float64 and np.allclose(out, g["out"], rtol=1e-12, atol=1e-9)
    m = g["geometry"][:, 4] >= 0
    assert np.allclose(E.restore_rectangle_rbox(g["origin"][m], g["geometry"][m]), g["out_pos_only"], rtol=1e-12, atol=1e-9)
    assert tuple(E.restore_rectangle_rbox(g["origin"][:0], g["geometry"][:0]).shape) == tuple(g["out_empty_shape"])


def test_order_points_golden(golden_dir):
    from oracle import decode as D
    from tensorflow_ocr_b200 import decode as P
    g = np.load(golden_dir + "/order_points.npz")
    for b, o, s in zip(g["boxes"], g["ordered"], g["sorted_poly"]):
        assert np.array_equal(D.order_points(b), o) and np.array_equal(P.order_points(b), o)
        assert np.array_equal(D.sort_poly(b), s) and np.array_equal(P.sort_poly(b), s)


def _loss_fp64(inp, M):
    """fp64 evaluation of nets/model.py:204-261 with the OHEM mask M held fixed."""
    def ce(logits, lab):
        x = logits.astype(np.float64)
        m = x.max(-1, keepdims=True)
        lse = np.log(np.exp(x - m).sum(-1)) + m[..., 0]
        return lse - np.take_along_axis(x, lab[..., None], -1)[..., 0]
    B = inp["pix_logits"].shape[0]
    M = M.reshape(B, -1).astype(np.float64)
    pl = inp["pix_lab"].reshape(B, -1).astype(np.int64)
    L_pix = (ce(inp["pix_logits"].reshape(B, -1, 2), pl) * M).sum() / (pl == 1).sum()
    ll = inp["link_lab"].reshape(B, -1, 8).astype(np.int64)
    lg = inp["link_logits"].reshape(B, -1, 8, 2)
    total = 0.0
    for d in range(8):
        c = ce(lg[:, :, d], ll[:, :, d])
        wp, wn = (ll[:, :, d] == 1) * M, (ll[:, :, d] == 0) * M
        total += (c * wp).sum() / wp.sum() + (c * wn).sum() / wn.sum()
    return total + 2 * L_pix


def test_gradients_match_finite_differences():
    """Analytic gradients of the oracle vs fp64 central differences (mask and counts held
    fixed, as TF autodiff does: they carry no gradient)."""
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import synth
    inp = synth.make_batch(21, 2, 8, 10, "C")
    base = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
    assert abs(_loss_fp64(inp, base["ohem_mask"]) - float(base["loss"])) < 1e-5
    rng = np.random.default_rng(0)
    for name, key in (("pix_logits", "grad_pixel"), ("link_logits", "grad_link")):
        for idx in rng.choice(inp[name].size, 24, replace=False):
            eps = 1e-3
            vals = []
            for sgn in (+1, -1):
                x = {k: v.astype(np.float64) for k, v in inp.items()}
                x[name].reshape(-1)[idx] += sgn * eps
                vals.append(_loss_fp64(x, base["ohem_mask"]))
            fd = (vals[0] - vals[1]) / (2 * eps)
            an = float(base[key].reshape(-1)[idx])
            assert abs(fd - an) <= 1e-4 * abs(an) + 1e-7, (name, idx, fd, an)


def test_generate_rbox_matches_reference_execution(golden_dir):
    """oracle/labels.py vs tool/pixellink_fn.py generate_rbox executed by tests/golden/make_golden.py."""
    from oracle import labels as OL
    g = np.load(golden_dir + "/generate_rbox.npz")
    for ci in range(int(g["n_cases"])):
        score, link, show, pm = OL.generate_rbox(int(g["h%d" % ci]), int(g["w%d" % ci]), g["xs%d" % ci], g["ys%d" % ci],
                                                 g["bboxes%d" % ci], g["ignored%d" % ci])
        assert np.array_equal(score, g["score%d" % ci])
        assert np.array_equal(link, g["link%d" % ci])
        assert np.array_equal(show, g["show%d" % ci])
        assert link.sum() > 0
        assert np.array_equal(OL.link_labels_from_ids(pm), OL.link_labels_from_ids_loop(pm))


import pytest


@pytest.mark.parametrize("tag", ["fast", "full"])
def test_link_graph_golden(golden_dir, tag):
    """D2/D3 pin: the reference scripts' own graph + DFS + box lines (test_pixellink_fast.py:111-202,
    test_pixellink.py:113-217), executed by line range on symmetric-link maps, against the canonical
    weakly-connected-component labelling and the cv2 boxes of the oracle."""
    from oracle import decode as D
    from util import link_graph_case
    c = link_graph_case(golden_dir, tag)
    assert c["n_groups"] >= 8
    P, L = D.thresholds(c["pix_logits"], c["link_logits"])
    assert np.array_equal(P, c["P"]) and np.array_equal(L, c["L"])
    labels, roots, sizes = D.link_components(c["P"], c["L"], c["min_size"])
    assert np.array_equal(labels, c["labels"])
    assert len(roots) == c["n_groups"]
    boxes, _ = D.component_boxes(labels, roots, c["scale"])
    assert np.array_equal(boxes, c["boxes"])
    # some components must have been dropped by the size filter, or the filter is not exercised
    all_labels, all_roots, _ = D.link_components(c["P"], c["L"], 0)
    assert len(all_roots) > len(roots)


def test_result_txt_golden(golden_dir, tmp_path):
    """N4: the ICDAR result file exactly as test_pixellink_fast.py:209-217 writes it."""
    from util import link_graph_case
    from tensorflow_ocr_b200.decode import write_result_txt
    c = link_graph_case(golden_dir, "fast")
    # the script writes boxes in gid (first-seed) order; the golden keeps them in label order, and the
    # file bytes in the script's order: compare as sets of lines
    write_result_txt(str(tmp_path / "res.txt"), c["boxes"])
    mine = open(tmp_path / "res.txt", "rb").read()
    assert mine.endswith(b"\r\n") and sorted(mine.split(b"\r\n")) == sorted(c["res_txt"].split(b"\r\n"))


def test_icdar_generate_rbox_golden(golden_dir):
    """N1, the EAST-fork generator (datasets/icdar.py:83-105, 486-539, executed by make_golden.py): the oracle's
    two-map restatement (last / first covering polygon) reproduces the order-dependent loop exactly."""
    from oracle import labels as OL
    g = np.load(golden_dir + "/icdar_generate_rbox.npz")
    for ci in range(int(g["n_cases"])):
        s = int(g["size%d" % ci])
        score, geo, tm = OL.icdar_generate_rbox((s, s), g["polys%d" % ci], g["tags%d" % ci])
        assert np.array_equal(score, g["score%d" % ci])
        assert np.array_equal(geo, g["geo%d" % ci])
        assert np.array_equal(tm, g["tmask%d" % ci])
        assert geo.sum() > 1000 and (geo[:, 0].sum() > 0 or geo[0].sum() > 0)   # the wrap-around borders are exercised


def test_head_logits_golden(golden_dir):
    """N3: oracle/head_logits.py vs the reference's own fusion lines executed (nets/pixellink.py:57-67,
    nets/model.py:129-141; make_golden.py::golden_head_logits)."""
    from oracle import head_logits as OH
    g = np.load(golden_dir + "/head_logits.npz")
    ep = {k: g["pl_" + k] for k in ("fc7", "conv5_3", "conv4_3", "conv3_3")}
    scopes = ["stage_%d_%s_fuse" % (st, kind) for kind in ("pixel", "link") for st in (6, 5, 4, 3)] + ["text_predication", "link_predication"]
    p = {s: (g["pl_w_" + s], g["pl_b_" + s]) for s in scopes}
    pix, link = OH.pixellink_layers(ep, p)
    assert pix.shape == g["pl_pixel_cls"].shape and link.shape[-1] == 16
    np.testing.assert_allclose(pix, g["pl_pixel_cls"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(link, g["pl_link_cls"], rtol=0, atol=2e-6)
    fm = [g["md_f%d" % i] for i in range(4)]
    q = {kind: [(g["md_%s_f%d_w" % (kind, i)], g["md_%s_f%d_scale" % (kind, i)], g["md_%s_f%d_shift" % (kind, i)]) for i in range(4)]
         + [(g["md_%s_out_w" % kind], None, g["md_%s_out_b" % kind])] for kind in ("pixel", "link")}
    pix, link = OH.model_head(fm, q)
    np.testing.assert_allclose(pix, g["md_pixel_4"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(link, g["md_link_4"], rtol=0, atol=2e-6)
    # unpool: TF's align_corners=False arithmetic on a ramp (x2: every other output sits on a source pixel,
    # the ones between are midpoints, the last row / column repeats)
    r = OH.resize_bilinear_x2(np.arange(4.0).reshape(1, 1, 4, 1))
    np.testing.assert_array_equal(r[0, 0, :, 0], [0, 0.5, 1, 1.5, 2, 2.5, 3, 3])


def test_resize_bilinear_x2_properties():
    """tf.image.resize_bilinear(align_corners=False) for an exact factor 2: constants stay constant, every other
    output sits on a source pixel, the ones between are midpoints, the last row / column repeats."""
    from oracle import head_logits as OH
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 5, 7, 3))
    y = OH.resize_bilinear_x2(x)
    assert y.shape == (2, 10, 14, 3)
    np.testing.assert_array_equal(y[:, ::2, ::2], x)
    np.testing.assert_allclose(y[:, 1:-1:2, ::2], 0.5 * (x[:, :-1] + x[:, 1:]), rtol=0, atol=1e-15)
    np.testing.assert_allclose(y[:, ::2, 1:-1:2], 0.5 * (x[:, :, :-1] + x[:, :, 1:]), rtol=0, atol=1e-15)
    np.testing.assert_array_equal(y[:, -1], y[:, -2])
    np.testing.assert_array_equal(y[:, :, -1], y[:, :, -2])
    np.testing.assert_array_equal(OH.resize_bilinear_x2(np.full((1, 3, 4, 2), 1.25)), np.full((1, 6, 8, 2), 1.25))
