"""Parity of the CUDA loss pipeline (through the C ABI) against the oracle and the goldens."""
import numpy as np
import pytest

from util import TOL, grad_close, rel_err

pytestmark = pytest.mark.gpu


def _run(inp, cfg=None, want_mask=True):
    import torch
    from tensorflow_ocr_b200 import head
    dev = torch.device("cuda", 0)
    t = {k: torch.as_tensor(inp[k]).to(dev) for k in ("pix_logits", "link_logits", "pix_lab", "link_lab")}
    out = head.pixellink_loss_raw(t["pix_logits"], t["link_logits"], t["pix_lab"], t["link_lab"],
                                  cfg or head.LossConfig(), want_grad=True, want_mask=want_mask)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _compare(out, ref, B, check_mask=True):
    from tensorflow_ocr_b200 import _lib
    st = out["stats"]
    assert rel_err(st[_lib.ST_TOTAL], ref["loss"]) <= TOL
    assert rel_err(out["grad_pixel"], ref["grad_pixel"]) <= TOL
    assert rel_err(out["grad_link"], ref["grad_link"]) <= TOL
    if check_mask:
        assert np.array_equal(out["ohem_mask"].astype(np.float32), ref["ohem_mask"]), "OHEM mask not bit-exact"
    if "w_pixel" in ref:   # elementwise, not only in the max norm
        grad_close(out["grad_pixel"], ref["grad_pixel"], ref["w_pixel"])
        grad_close(out["grad_link"], ref["grad_link"], ref["w_link"])
    if "sum_wp" in ref:
        assert np.array_equal(st[_lib.ST_SUM_WP:_lib.ST_SUM_WP + 8], ref["sum_wp"])
        assert np.array_equal(st[_lib.ST_SUM_WN:_lib.ST_SUM_WN + 8], ref["sum_wn"])
        assert st[_lib.ST_N_SEG_POS] == ref["n_seg_pos"]
        thr = st[_lib.ST_THR:_lib.ST_THR + B]
        assert np.array_equal(np.isnan(thr), np.isnan(ref["thr"]))
        ok = ~np.isnan(thr)
        # the threshold VALUE may differ in the last bits (CUDA expf vs numpy exp); the mask is what is bit-exact
        assert np.allclose(thr[ok], ref["thr"][ok], rtol=1e-6, atol=0), "OHEM thresholds differ"


def test_model_loss_golden(golden_dir, cuda_dev):
    """nets/model.py loss executed by the reference source (B=14) vs the CUDA path."""
    g = np.load(golden_dir + "/model_loss_b14.npz")
    out = _run(g)
    from tensorflow_ocr_b200 import _lib
    assert rel_err(out["stats"][_lib.ST_TOTAL], g["loss"]) <= TOL
    assert rel_err(out["grad_pixel"], g["grad_pixel"]) <= TOL
    assert rel_err(out["grad_link"], g["grad_link"]) <= TOL
    assert np.array_equal(out["ohem_mask"].astype(np.float32), g["ohem_mask"])


def test_model_loss_golden_nan(golden_dir, cuda_dev):
    g = np.load(golden_dir + "/model_loss_b14_nopos.npz")
    out = _run(g)
    assert np.isnan(out["stats"][0]) and np.isnan(g["loss"])
    assert rel_err(out["grad_pixel"], g["grad_pixel"]) <= TOL
    assert np.isnan(out["grad_link"]).all() and np.isnan(g["grad_link"]).all()


@pytest.mark.parametrize("B,H,W,edge", [(32, 128, 128, True), (1, 128, 128, False), (5, 24, 40, True), (8, 64, 64, True), (3, 37, 53, True),
                                        (2, 240, 240, False), (4, 192, 192, True), (2, 96, 200, False)])
def test_model_loss_vs_oracle(B, H, W, edge, cuda_dev):
    """(32, 128, 128) is the headline shape of BASELINE config 2 — batch-wide normalisers over 32 images and the
    8-keys-per-thread tier of the cluster selection kernel, which is what bench.py launches."""
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import synth
    inp = synth.make_batch(2, B, H, W, "G", edge_images=edge)
    ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
    _compare(_run(inp), ref, B)


def test_vgg16_ohem_loss(golden_dir, cuda_dev):
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import _lib, head, synth
    cfg = head.LossConfig(variant=_lib.VARIANT_POS_ONLY)
    g = np.load(golden_dir + "/vgg16_ohem_loss.npz")
    out = _run(g, cfg)
    assert rel_err(out["stats"][0], g["loss"]) <= TOL
    assert rel_err(out["grad_pixel"], g["grad_pixel"]) <= TOL
    assert rel_err(out["grad_link"], g["grad_link"]) <= TOL
    inp = synth.make_batch(3, 6, 48, 32, "G", edge_images=True)
    ref = O.ohem_loss_vgg16(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
    _compare(_run(inp, cfg), ref, 6)


def test_pixellink_build_loss(golden_dir, cuda_dev):
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import _lib, head, synth
    cfg = head.LossConfig(variant=_lib.VARIANT_PIXELLINK)
    g = np.load(golden_dir + "/pixellink_build_loss.npz")
    out = _run(g, cfg)
    st = out["stats"]
    assert rel_err(2 * st[_lib.ST_L_PIX], g["losses"][0]) <= TOL
    assert rel_err(st[_lib.ST_LINK_TOTAL], g["losses"][1]) <= TOL
    assert rel_err(out["grad_pixel"], g["grad_pixel"]) <= TOL
    assert rel_err(out["grad_link"], g["grad_link"]) <= TOL
    assert np.array_equal(out["ohem_mask"].astype(np.float32), g["ohem_mask"])
    inp = synth.make_batch(4, 7, 40, 56, "G", edge_images=True)
    ref = O.build_loss_pixellink(inp["pix_logits"], inp["link_logits"], inp["pix_lab"], inp["link_lab"])
    out = _run(inp, cfg)
    assert rel_err(out["stats"][0], ref["loss"]) <= TOL
    assert rel_err(out["grad_pixel"], ref["grad_pixel"]) <= TOL
    assert rel_err(out["grad_link"], ref["grad_link"]) <= TOL
    assert np.array_equal(out["ohem_mask"].astype(np.float32), ref["ohem_mask"])
    assert out["stats"][_lib.ST_N_SEG_POS] == ref["n_seg_pos"]


def test_focal_loss_vs_oracle(cuda_dev):
    """L10 (parity unpinned: focal loss is not in the reference)."""
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import _lib, head, synth
    inp = synth.make_batch(5, 6, 48, 48, "G", edge_images=True)
    ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"], term="focal")
    _compare(_run(inp, head.LossConfig(term=_lib.TERM_FOCAL)), ref, 6)


def test_ohnm_batch_api(cuda_dev):
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import synth
    from tensorflow_ocr_b200.nets import model
    inp = synth.make_batch(6, 4, 32, 32, "G", edge_images=True)
    B = 4
    scores = O.softmax2(inp["pix_logits"].reshape(B, -1, 2))[:, :, 0]
    lab = inp["pix_lab"].reshape(B, -1).astype(np.int32)
    pos, neg = lab == 1, lab == 0
    ref, _ = O.OHNM_batch(14, scores, pos, neg)
    got = model.OHNM_batch(14, scores, pos, neg)
    assert np.array_equal(got, ref)
    for b in range(B):
        r1, _ = O.OHNM_single_image(scores[b], int(pos[b].sum()), neg[b])
        g1 = model.OHNM_single_image(scores[b], int(pos[b].sum()), neg[b])
        assert np.array_equal(g1, r1)


def test_autograd_drop_in(cuda_dev):
    import torch
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import synth
    from tensorflow_ocr_b200.nets import model
    inp = synth.make_batch(8, 3, 32, 32, "G")
    ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
    yp = torch.tensor(inp["pix_logits"], device=cuda_dev, requires_grad=True)
    yl = torch.tensor(inp["link_logits"], device=cuda_dev, requires_grad=True)
    l = model.loss(torch.tensor(inp["pix_lab"], device=cuda_dev), yp, torch.tensor(inp["link_lab"], device=cuda_dev), yl,
                   torch.ones(3, 32, 32, 1, device=cuda_dev))
    (3.0 * l).backward()
    assert rel_err(l.item(), ref["loss"]) <= TOL
    assert rel_err(yp.grad.cpu().numpy(), 3.0 * ref["grad_pixel"]) <= TOL
    assert rel_err(yl.grad.cpu().numpy(), 3.0 * ref["grad_link"]) <= TOL
    # numpy in -> numpy scalar out (py_func style)
    v = model.loss(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"], inp["training_mask"])
    assert rel_err(v, ref["loss"]) <= TOL


def test_full_size_properties(cuda_dev):
    """BASELINE config 2 size (B=32, 128x128), size-independent properties:
    mask count identity, gradient sums to zero per 2-way pair, determinism."""
    from tensorflow_ocr_b200 import _lib, synth
    inp = synth.make_batch(2, 32, 128, 128, "G", edge_images=True)
    o1 = _run(inp)
    o2 = _run(inp)
    for k in ("stats", "grad_pixel", "grad_link", "ohem_mask"):
        assert np.array_equal(o1[k], o2[k], equal_nan=True), "non-deterministic " + k
    st = o1["stats"]
    assert st[_lib.ST_N_SELECTED] == o1["ohem_mask"].sum()
    gp = o1["grad_pixel"]
    assert np.array_equal(gp[..., 0], -gp[..., 1])
    gl = o1["grad_link"].reshape(32, 128, 128, 8, 2)
    assert np.array_equal(gl[..., 0], -gl[..., 1])
    # unselected pixels carry no gradient
    assert np.all(gp[o1["ohem_mask"] == 0] == 0)
    # per image: #selected negatives >= min(3 n_pos, n_neg), equality unless ties at the threshold
    lab = inp["pix_lab"].reshape(32, -1).astype(np.int32)
    m = o1["ohem_mask"].reshape(32, -1)
    for b in range(32):
        npos, nneg = int((lab[b] == 1).sum()), int((lab[b] == 0).sum())
        nsel = int((m[b] == 1).sum()) - npos
        k = min(3 * npos, nneg) if npos > 0 else 0
        assert nsel >= k and (k > 0 or nsel == 0)


@pytest.mark.parametrize("seed", range(5))
def test_loss_random_shapes_and_variants(seed, cuda_dev):
    """Ragged sizes x all three weight variants x both terms against the oracle."""
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import _lib, head, synth
    rng = np.random.default_rng(200 + seed)
    B, H, W = int(rng.integers(1, 7)), int(rng.integers(2, 90)), int(rng.integers(2, 120))
    inp = synth.make_batch(40 + seed, B, H, W, "G", edge_images=(B >= 4))
    ratio = int(rng.integers(1, 5))
    ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"], ratio=ratio)
    _compare(_run(inp, head.LossConfig(neg_pos_ratio=ratio)), ref, B)
    ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"], weight_mode="pos_only",
                       term="focal")
    _compare(_run(inp, head.LossConfig(variant=_lib.VARIANT_POS_ONLY, term=_lib.TERM_FOCAL)), ref, B)
    ref = O.build_loss_pixellink(inp["pix_logits"], inp["link_logits"], inp["pix_lab"], inp["link_lab"], ratio)
    out = _run(inp, head.LossConfig(variant=_lib.VARIANT_PIXELLINK, neg_pos_ratio=ratio))
    assert rel_err(out["stats"][0], ref["loss"]) <= TOL
    assert rel_err(out["grad_pixel"], ref["grad_pixel"]) <= TOL
    assert rel_err(out["grad_link"], ref["grad_link"]) <= TOL
    assert np.array_equal(out["ohem_mask"].astype(np.float32), ref["ohem_mask"])


def test_loss_large_image_general_select_path(cuda_dev):
    """> 65536 px per image: K0 + the one-CTA-per-image selection (keys not register resident)."""
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import head, synth
    for H, W in ((272, 256), (512, 160)):
        inp = synth.make_batch(77, 2, H, W, "G", edge_images=False)
        ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
        _compare(_run(inp, head.LossConfig()), ref, 2)


@pytest.mark.parametrize("variant", ["model", "pixellink"])
def test_split_counts_hint_is_result_neutral(variant, cuda_dev):
    """plh_loss_params.reserved[0] bit 1 only moves the mask + normaliser pass into its own kernel."""
    from tensorflow_ocr_b200 import _lib, head, synth
    inp = synth.make_batch(91, 5, 48, 80, "G", edge_images=True)
    v = _lib.VARIANT_MODEL if variant == "model" else _lib.VARIANT_PIXELLINK
    a = _run(inp, head.LossConfig(variant=v))
    b = _run(inp, head.LossConfig(variant=v, split_counts=True))
    for k in ("stats", "grad_pixel", "grad_link", "ohem_mask"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("hw", [(256, 256), (257, 255), (64, 32), (8, 256)])
def test_loss_select_path_boundaries(hw, cuda_dev):
    """Image sizes at the edges of the cluster kernel's register tiers (2 / 8 / 32 keys per thread: 4096,
    16384, 65536 px) and one pixel beyond the last tier (general path)."""
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import head, synth
    H, W = hw
    inp = synth.make_batch(123 + H, 2, H, W, "G", edge_images=False)
    ref = O.loss_model(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
    _compare(_run(inp, head.LossConfig()), ref, 2)


def test_ohnm_rejects_scores_outside_unit_interval(cuda_dev):
    """The standalone OHNM ranks fp32 bit patterns: only probabilities in [0, 1] are ordered by them."""
    from tensorflow_ocr_b200.nets import model
    sc = np.random.default_rng(0).uniform(size=(2, 64)).astype(np.float32)
    pos = np.zeros((2, 64), bool)
    pos[:, :5] = True
    neg = ~pos
    model.OHNM_batch(14, sc, pos, neg)
    for bad in (-0.25, 1.5, np.nan):
        s2 = sc.copy()
        s2[1, 7] = bad
        with pytest.raises(ValueError):
            model.OHNM_batch(14, s2, pos, neg)
        with pytest.raises(ValueError):
            model.OHNM_single_image(s2[1], 5, neg[1])


def test_loss_terms_at_config5_shape(cuda_dev):
    """BASELINE config 5's per-GPU shape (32 x 192x192: the 32-keys-per-thread tier of the selection kernel, inputs
    larger than half of L2): CE+OHEM, focal (parity unpinned) and the dice head against the oracle, gradients
    elementwise."""
    import torch
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import _lib, head, synth
    B, H, W = 32, 192, 192
    inp = synth.make_batch(5, B, H, W, "G", edge_images=True)
    a = (inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"])
    _compare(_run(inp), O.loss_model(*a), B)
    _compare(_run(inp, head.LossConfig(term=_lib.TERM_FOCAL)), O.loss_model(*a, term="focal"), B)
    pl, ll = inp["pix_logits"], inp["link_logits"].reshape(B, H, W, 8, 2)
    pp = (1.0 / (1.0 + np.exp(-(pl[..., 1:2] - pl[..., 0:1])))).astype(np.float32)
    lp = (1.0 / (1.0 + np.exp(-(ll[..., 1] - ll[..., 0])))).astype(np.float32)
    m = (np.random.default_rng(1).uniform(size=(B, H, W, 1)) > 0.1).astype(np.float32)
    ref = O.loss_vgg16_dice(inp["pix_lab"], pp, inp["link_lab"], lp, m)
    t = lambda x: torch.as_tensor(x).to(cuda_dev)
    outv, gp, gl = head.dice_head_raw(t(inp["pix_lab"]), t(pp), t(inp["link_lab"]), t(lp), t(m))
    torch.cuda.synchronize()
    assert rel_err(outv[0].item(), ref["loss"]) <= TOL
    assert rel_err(gp.cpu().numpy(), ref["grad_pixel"]) <= TOL
    assert rel_err(gl.cpu().numpy(), ref["grad_link"]) <= TOL
