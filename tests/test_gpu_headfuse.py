"""N3 on the GPU: the logit producer (csrc/headfuse.cu through the C ABI and the nets mirrors) against the
reference-executed golden and the float64 oracle.

Tolerance (the contract of include/plhead.h): |gpu - f64| <= 1e-5 * max|f64| per tensor — fp32 FMA accumulation of K up
to 2048 products."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err <= tol, err
    return err


def _pl_params(g=None, rng=None, chans=None):
    scopes = [("stage_%d_%s_fuse" % (st, kind), name, n) for kind, n in (("pixel", 2), ("link", 16))
              for st, name in ((6, "fc7"), (5, "conv5_3"), (4, "conv4_3"), (3, "conv3_3"))]
    p = {}
    for scope, name, n in scopes:
        p[scope] = (g["pl_w_" + scope], g["pl_b_" + scope]) if g is not None else (
            (rng.standard_normal((chans[name], n)) / np.sqrt(chans[name])).astype(np.float32), (0.1 * rng.standard_normal(n)).astype(np.float32))
    for scope, n in (("text_predication", 2), ("link_predication", 16)):
        p[scope] = (g["pl_w_" + scope], g["pl_b_" + scope]) if g is not None else (
            (rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32), (0.1 * rng.standard_normal(n)).astype(np.float32))
    return p


def test_pixellink_layers_golden(golden_dir, cuda_dev):
    from tensorflow_ocr_b200.nets import pixellink
    g = np.load(golden_dir + "/head_logits.npz")
    ep = {k: g["pl_" + k] for k in ("fc7", "conv5_3", "conv4_3", "conv3_3")}
    pix, link = pixellink.pixellink_layers(ep, _pl_params(g))
    assert pix.dtype == np.float32 and pix.shape == g["pl_pixel_cls"].shape and link.shape == g["pl_link_cls"].shape
    _close(pix, g["pl_pixel_cls"])
    _close(link, g["pl_link_cls"])


def test_model_feature_fusion_golden(golden_dir, cuda_dev):
    from tensorflow_ocr_b200.nets import model
    g = np.load(golden_dir + "/head_logits.npz")
    fm = [g["md_f%d" % i] for i in range(4)]
    q = {kind: [(g["md_%s_f%d_w" % (kind, i)], g["md_%s_f%d_scale" % (kind, i)], g["md_%s_f%d_shift" % (kind, i)]) for i in range(4)]
         + [(g["md_%s_out_w" % kind], None, g["md_%s_out_b" % kind])] for kind in ("pixel", "link")}
    pix, link = model.feature_fusion(fm, q)
    _close(pix, g["md_pixel_4"])
    _close(link, g["md_link_4"])


@pytest.mark.parametrize("B,H,W,chans", [
    (2, 32, 48, {"fc7": 1024, "conv5_3": 512, "conv4_3": 512, "conv3_3": 256}),   # VGG-16 widths
    (1, 20, 28, {"fc7": 40, "conv5_3": 36, "conv4_3": 68, "conv3_3": 100}),      # K not a multiple of the 32-channel chunk, ragged tiles
    (3, 8, 8, {"fc7": 8, "conv5_3": 4, "conv4_3": 12, "conv3_3": 4}),            # fewer channels than one chunk, fewer pixels than one tile
])
def test_pixellink_layers_vs_oracle(B, H, W, chans, cuda_dev):
    import torch
    from oracle import head_logits as OH
    from tensorflow_ocr_b200.nets import pixellink
    rng = np.random.default_rng(B * 100 + H)
    ep = {"fc7": rng.standard_normal((B, H // 4, W // 4, chans["fc7"])), "conv5_3": rng.standard_normal((B, H // 4, W // 4, chans["conv5_3"])),
          "conv4_3": rng.standard_normal((B, H // 2, W // 2, chans["conv4_3"])), "conv3_3": rng.standard_normal((B, H, W, chans["conv3_3"]))}
    ep = {k: v.astype(np.float32) for k, v in ep.items()}
    p = _pl_params(rng=rng, chans=chans)
    o_pix, o_link = OH.pixellink_layers(ep, p)
    pix, link = pixellink.pixellink_layers({k: torch.as_tensor(v).to(cuda_dev) for k, v in ep.items()}, p)
    assert pix.is_cuda and tuple(pix.shape) == (B, H, W, 2) and tuple(link.shape) == (B, H, W, 16)
    _close(pix.cpu().numpy(), o_pix)
    _close(link.cpu().numpy(), o_link)


def test_feature_fusion_vs_oracle_resnet_widths(cuda_dev):
    from oracle import head_logits as OH
    from tensorflow_ocr_b200.nets import model
    rng = np.random.default_rng(9)
    B, H, W, fch = 1, 32, 32, [2048, 1024, 512, 256]
    fm = [rng.standard_normal((B, H >> (3 - i), W >> (3 - i), fch[i])).astype(np.float32) for i in range(4)]
    q = {}
    for kind, n in (("pixel", 2), ("link", 16)):
        q[kind] = [((rng.standard_normal((fch[i], n)) / np.sqrt(fch[i])).astype(np.float32), rng.uniform(0.5, 1.5, n).astype(np.float32),
                    (0.3 * rng.standard_normal(n)).astype(np.float32)) for i in range(4)]
        q[kind].append(((rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32), None, (0.1 * rng.standard_normal(n)).astype(np.float32)))
    o_pix, o_link = OH.model_head(fm, q)
    pix, link = model.feature_fusion(fm, q)
    _close(pix, o_pix)
    _close(link, o_link)


def test_produced_logits_feed_the_loss(cuda_dev):
    """The producer's two tensors are the loss / decode entry points' inputs as they are (layout, dtype, alignment)."""
    import torch
    from tensorflow_ocr_b200 import head, synth
    from tensorflow_ocr_b200.nets import pixellink
    rng = np.random.default_rng(4)
    B, H, W = 2, 32, 32
    chans = {"fc7": 64, "conv5_3": 32, "conv4_3": 32, "conv3_3": 32}
    ep = {"fc7": rng.standard_normal((B, H // 4, W // 4, 64)), "conv5_3": rng.standard_normal((B, H // 4, W // 4, 32)),
          "conv4_3": rng.standard_normal((B, H // 2, W // 2, 32)), "conv3_3": rng.standard_normal((B, H, W, 32))}
    pix, link = pixellink.pixellink_layers({k: torch.as_tensor(v.astype(np.float32)).to(cuda_dev) for k, v in ep.items()},
                                           _pl_params(rng=rng, chans=chans))
    lab = synth.make_batch(2, B, H, W, "G")
    out = head.loss_and_decode_raw(pix, link, torch.as_tensor(lab["pix_lab"]).to(cuda_dev), torch.as_tensor(lab["link_lab"]).to(cuda_dev))
    torch.cuda.synchronize()
    assert np.isfinite(out["stats"].cpu().numpy()).all()


def test_threshold_words_from_the_producer(cuda_dev):
    """The last level can emit the decode's threshold words: identical to plh_decode_flags on the logits it wrote,
    and the decode started from them gives the boxes of the decode started from the logits."""
    import torch
    from tensorflow_ocr_b200 import head
    from tensorflow_ocr_b200.nets import pixellink
    rng = np.random.default_rng(12)
    B, H, W = 2, 64, 64
    chans = {"fc7": 64, "conv5_3": 32, "conv4_3": 32, "conv3_3": 32}
    ep = {"fc7": rng.standard_normal((B, H // 4, W // 4, 64)), "conv5_3": rng.standard_normal((B, H // 4, W // 4, 32)),
          "conv4_3": rng.standard_normal((B, H // 2, W // 2, 32)), "conv3_3": rng.standard_normal((B, H, W, 32))}
    ep = {k: torch.as_tensor(v.astype(np.float32)).to(cuda_dev) for k, v in ep.items()}
    cfg = head.DecodeConfig(pixel_thresh=0.6, link_thresh=0.55, min_size=3, max_boxes=256)
    pix, link, flags = pixellink.pixellink_layers(ep, _pl_params(rng=rng, chans=chans), decode_config=cfg)
    ref = head.decode_flags_raw(pix, link, cfg)["flags"]
    assert torch.equal(flags, ref) and int((flags != 0).sum()) > 0
    a = head.decode_from_flags_raw(flags, cfg, {}, True)
    b = head.decode_raw(pix, link, cfg, {}, True)
    assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["n_boxes"], b["n_boxes"])
    for i, n in enumerate(a["n_boxes"].tolist()):      # rows past n_boxes are not written
        assert n > 0 and torch.equal(a["boxes"][i, :n], b["boxes"][i, :n]), i


def test_argument_errors(cuda_dev):
    import torch
    from tensorflow_ocr_b200 import head
    x = torch.zeros((1, 8, 8, 6), device=cuda_dev)
    w = torch.zeros((6, 18), device=cuda_dev)
    with pytest.raises(ValueError):
        head.head_fuse_level_raw([(x, w, None, None, False)])                       # K % 4 != 0
    x = torch.zeros((1, 7, 8, 8), device=cuda_dev)
    w = torch.zeros((8, 18), device=cuda_dev)
    with pytest.raises(ValueError):
        head.head_fuse_level_raw([(x, w, None, None, False)], prev=torch.zeros((1, 3, 4, 18), device=cuda_dev))   # odd H
    with pytest.raises(ValueError):
        head.head_fuse_level_raw([(x, torch.zeros((8, 16), device=cuda_dev), None, None, False)])                 # 18 columns
