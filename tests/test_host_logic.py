"""CPU: host-side logic, the C-ABI library's exports, and the product/oracle separation."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    """Every `PLH_API ... plh_*(` in include/plhead.h must be exported by libplhead.so and bound
    by tensorflow_ocr_b200/_lib.py (no compute calls here: there is no GPU)."""
    from tensorflow_ocr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    hdr = open(os.path.join(ROOT, "include", "plhead.h")).read()
    declared = sorted(set(re.findall(r"PLH_API\s+[\w\s\*]+?\b(plh_\w+)\s*\(", hdr)))
    assert len(declared) >= 14
    assert sorted(_lib.SIGNATURES) == declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    l = _lib.load()
    assert l.plh_version() == 100
    assert l.plh_strerror(-4).decode().startswith("workspace")
    assert l.plh_workspace_bytes(_lib.OP_LOSS, 32, 128, 128, 0) > 32 * 128 * 128 * 5
    assert l.plh_workspace_bytes(_lib.OP_LOSS, 0, 128, 128, 0) == 0


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call: error codes are checkable on CPU."""
    from tensorflow_ocr_b200 import _lib
    l = _lib.load()
    lp = _lib.LossParams(0, 0, 3, 0.25, 2.0)
    rc = l.plh_pixellink_loss(None, None, None, None, None, 1, 8, 8, ctypes.byref(lp), None, None, None, None, None,
                              None, None, 0, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc, "x")
    buf = (ctypes.c_float * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    lp = _lib.LossParams(7, 0, 3, 0.25, 2.0)
    assert l.plh_pixellink_loss(p, p, p, p, None, 1, 4, 4, ctypes.byref(lp), p, None, None, None, None, None, p, 0,
                                None) == -5
    lp = _lib.LossParams(0, 0, 3, 0.25, 2.0)
    assert l.plh_pixellink_loss(p, p, p, p, None, 1, 4, 4, ctypes.byref(lp), p, None, None, None, None, None, p, 16,
                                None) == -4
    assert l.plh_pixellink_loss(p, p, p, p, None, 0, 4, 4, ctypes.byref(lp), p, None, None, None, None, None, p, 16,
                                None) == -2
    dp = _lib.DecodeParams(0.8, 0.9, 10, 16, 0.5, 3.75)  # scale < 1 is refused
    assert l.plh_decode(p, p, 1, 4, 4, ctypes.byref(dp), p, p, p, None, None, p, 1 << 20, None) == -5


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tensorflow_ocr_b200.nets import model
    from tensorflow_ocr_b200 import synth
    inp = synth.make_batch(1, 1, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.loss(inp["pix_lab"], inp["pix_logits"], inp["link_lab"], inp["link_logits"], inp["training_mask"])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tensorflow_ocr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


def test_drop_in_signatures_match_the_reference_names():
    """Same names, same positional parameters as the reference."""
    import inspect
    from tensorflow_ocr_b200.datasets import icdar
    from tensorflow_ocr_b200.nets import model, model_vgg_16, pixellink
    from tensorflow_ocr_b200.tool import pixellink_fn
    from tensorflow_ocr_b200 import decode

    def params(fn):
        return list(inspect.signature(fn).parameters)

    five = ["y_true_pixel", "y_pred_pixel", "y_true_link", "y_pred_link", "training_mask"]
    assert params(model.loss) == five
    assert params(model_vgg_16.loss) == five
    assert params(model_vgg_16.ohem_loss) == five
    assert params(model.dice_coefficient) == ["y_true_cls", "y_pred_cls", "training_mask"]
    assert params(model.OHNM_single_image) == ["scores", "n_pos", "neg_mask"]
    assert params(model.OHNM_batch) == ["batch_size", "neg_conf", "pos_mask", "neg_mask"]
    assert params(model.get_pos_and_neg_masks) == ["labels"]
    assert params(model_vgg_16.cal_link_loss) == ["link_gt", "link_pred", "W_pixel"]
    assert params(pixellink.PixelLinkNet.build_loss) == ["self", "pixel_labels", "link_labels", "do_summary"]
    sig = inspect.signature(pixellink_fn.pixel_detect)
    assert list(sig.parameters) == ["score_map", "geo_map", "score_map_thresh", "link_thresh"]
    assert sig.parameters["score_map_thresh"].default == 0.8 and sig.parameters["link_thresh"].default == 0.8
    assert params(pixellink_fn.tf_pixel_detect) == ["score_map", "geo_map", "score_map_thresh", "link_thresh"]
    assert params(pixellink_fn.generate_rbox) == ["h", "w", "xs", "ys", "bboxes", "ignored"]
    assert params(pixellink_fn.tf_pixellink_get_rbox) == ["img_size", "xs", "ys", "bboxes", "ignored"]
    assert params(icdar.restore_rectangle) == ["origin", "geometry"]
    sig = inspect.signature(decode.decode_pixellink)
    assert sig.parameters["pixel_thresh"].default == 0.8 and sig.parameters["link_thresh"].default == 0.9
    assert sig.parameters["min_size"].default == 10 and sig.parameters["scale"].default == (4.0, 3.75)


def test_synth_is_deterministic_and_tie_robust():
    from tensorflow_ocr_b200 import synth
    a = synth.make_batch(2, 3, 32, 32, "G", edge_images=True)
    b = synth.make_batch(2, 3, 32, 32, "G", edge_images=True)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["pix_logits"] * 128, np.round(a["pix_logits"] * 128))   # on the 1/128 grid
    assert a["pix_lab"][2].sum() == 0 and a["pix_lab"][0].all()                        # no_pos last, all_pos first of the 3
    # link labels: text pixels on the border link everywhere (tool/pixellink_fn.py:9-11)
    allpos = a["link_lab"][0]
    assert allpos[0, :, :].all() and allpos[:, 0, :].all()
    s = synth.make_image(1, 0, 16, 16, "S")
    z = (s["link_logits"].reshape(16, 16, 8, 2)[..., 1] - s["link_logits"].reshape(16, 16, 8, 2)[..., 0])
    assert np.array_equal(z[5, 5, 3], z[5, 6, 0]) and np.array_equal(z[5, 5, 7], z[6, 5, 6])


def test_shard_bounds_are_a_contiguous_partition():
    from tensorflow_ocr_b200 import dist
    for B in (1, 7, 32, 64, 256):
        for G in (1, 2, 4, 8):
            spans = [dist.shard_bounds(B, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            if B % G == 0:
                assert all(hi - lo == B // G for lo, hi in spans)      # == tf.split (multigpu_train.py:111-114)


def test_oracle_edge_cases():
    from oracle import pixellink_loss as O
    sc = np.array([0.9, 0.1, 0.5, 0.5, 0.7], np.float32)
    neg = np.array([1, 1, 1, 1, 0], bool)
    m, thr = O.OHNM_single_image(sc, 1, neg)             # k = min(3, 4) = 3 -> thr 0.5, ties all selected
    assert thr == np.float32(0.5) and m.tolist() == [0, 1, 1, 1, 0]
    m, thr = O.OHNM_single_image(sc, 0, neg)             # no positives -> nothing
    assert m.sum() == 0 and np.isnan(thr)
    m, thr = O.OHNM_single_image(sc, 2, np.zeros(5, bool))  # no negatives (TF would raise) -> nothing
    assert m.sum() == 0
    # pixellink.py variant: zeros of the non-negatives occupy the first n_pos slots (quirk Q5)
    m, thr = O.OHNM_single_image_pixellink(sc, 1, neg, 3)   # k = 3 over [0.9,0.1,0.5,0.5,0] -> thr 0.5
    assert thr == np.float32(0.5) and m.tolist() == [0, 1, 1, 1, 0]
    m, thr = O.OHNM_single_image_pixellink(sc, 1, neg, 1)   # k = 1 -> thr 0 -> only negatives with score 0
    assert thr == 0 and m.sum() == 0


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import pixellink_loss as O
    from tensorflow_ocr_b200 import _lib, synth
    from tensorflow_ocr_b200 import dist as pdist
    full = synth.make_batch(31, 4, 16, 16, "G")
    shard = pdist.shard_batch(full, world, rank)
    r = O.loss_model(shard["pix_lab"], shard["pix_logits"], shard["link_lab"], shard["link_logits"])
    stats = torch.zeros(_lib.STATS_FLOATS + 2)
    stats[_lib.ST_TOTAL] = float(r["loss"])
    stats[_lib.ST_N_SEG_POS] = float(r["n_seg_pos"])
    red = pdist.allreduce_loss_stats(stats)
    q.put((rank, float(r["loss"]), float(red[_lib.ST_TOTAL]), float(red[_lib.ST_N_SEG_POS]), float(r["n_seg_pos"])))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_scalar_allreduce():
    """N>1 host path on CPU (gloo, world_size 2): contiguous batch shards, shard-local
    normalisers, one all-reduce of the loss scalars (mean of tower losses, summed counts)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    (_, l0, m0, n0, p0), (_, l1, m1, n1, p1) = res
    assert abs(m0 - (l0 + l1) / 2) < 1e-6 and abs(m1 - m0) < 1e-9
    assert n0 == p0 + p1 == n1


def test_scheduling_hints_map_to_the_reserved_words():
    """LossConfig / DecodeConfig -> plh_*_params.reserved[0] bits documented in include/plhead.h."""
    import dataclasses
    from tensorflow_ocr_b200 import head
    assert head.LossConfig().c_struct().reserved[0] == 0
    assert head.LossConfig(main_only=True).c_struct().reserved[0] == 1
    assert head.LossConfig(split_counts=True).c_struct().reserved[0] == 2
    assert head.LossConfig(chain_pdl=True, main_only=True).c_struct().reserved[0] == 5
    d = head.DecodeConfig()
    assert d.c_struct().reserved[0] == 0
    assert dataclasses.replace(d, phase=4).c_struct().reserved[0] == 4
    assert dataclasses.replace(d, phase=8).c_struct().reserved[0] == 8
    hdr = open(os.path.join(ROOT, "include", "plhead.h")).read()
    for word in ("bit 2: only the first kernel", "bit 3: everything after it", "Bit 1: scheduling hint", "Bit 2: scheduling hint"):
        assert word in hdr, word
