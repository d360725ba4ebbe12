#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference; the GPU box does not have
it):  python tests/golden/make_golden.py

The reference cannot be imported as modules (Python 2 syntax elsewhere in the
files, TensorFlow 1.4 / tf.contrib.slim / `config` not installable).  So each
hot-path function is cut out of its file by name, and exec'd unmodified in a
namespace where ``tf`` / ``slim`` are tests/golden/tf_shim.py (TF op semantics
restated over torch CPU fp32, autograd for the gradients) and ``np`` / ``dist``
are the real numpy / scipy.  Pure-numpy functions (pixel_detect,
restore_rectangle_rbox, order_points, sort_poly) run as they are.

Nothing from the reference is copied into the repo: only the numeric
inputs/outputs are stored.
"""
from __future__ import annotations

import os
import re
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf_shim as tf  # noqa: E402
from tensorflow_ocr_b200 import synth  # noqa: E402

REF = os.environ.get("PLH_REFERENCE", "/root/reference")


def cut(path, name, indent=""):
    """Source of ``def name`` (at the given indent) from a reference file, dedented."""
    lines = open(os.path.join(REF, path)).read().split("\n")
    start = None
    for i, l in enumerate(lines):
        if re.match(r"^%sdef %s\(" % (indent, re.escape(name)), l):
            start = i
            break
    assert start is not None, (path, name)
    end = len(lines)
    for j in range(start + 1, len(lines)):
        l = lines[j]
        if l.strip() == "":
            continue
        cur = len(l) - len(l.lstrip())
        if cur <= len(indent) and not l.lstrip().startswith("#"):
            end = j
            break
    return textwrap.dedent("\n".join(lines[start:end])), (start + 1, end)


def ns_tf():
    return dict(tf=tf, slim=tf.slim, np=np, xrange=range)


def T(x, grad=False):
    t = torch.tensor(np.asarray(x), dtype=torch.float32)
    t.requires_grad_(grad)
    return t


def save(name, **arrs):
    out = os.path.join(HERE, name + ".npz")
    np.savez_compressed(out, **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", out, os.path.getsize(out), "bytes")


# ----------------------------------------------------------------------------- nets/model.py
def golden_model_loss():
    ns = ns_tf()
    spans = {}
    for fn in ("dice_coefficient", "OHNM_single_image", "OHNM_batch", "get_pos_and_neg_masks", "loss"):
        src, spans[fn] = cut("nets/model.py", fn)
        exec(src, ns)
    print("nets/model.py spans", spans)
    # The reference hard-codes OHNM_batch(14, ...) (model.py:220) -> B must be 14.
    B, H, W = 14, 16, 16
    kinds = ["normal"] * B
    kinds[12], kinds[13] = "no_pos", "many_pos"
    imgs = [synth.make_image(91, i, H, W, "G", kinds[i]) for i in range(B)]
    inp = {k: np.stack([im[k] for im in imgs]) for k in imgs[0]}
    yp, yl = T(inp["pix_logits"], True), T(inp["link_logits"], True)
    mask = torch.ones(B, H, W, 1)
    out = ns["loss"](T(inp["pix_lab"]), yp, T(inp["link_lab"]), yl, mask)
    out.backward()
    # side values: the OHEM mask as the reference computes it
    with torch.no_grad():
        pl = tf.cast(tf.reshape(T(inp["pix_lab"]), [B, -1]), tf.int32)
        sc = tf.slim.softmax(tf.reshape(yp, [B, -1, 2]))[:, :, 0]
        pos, neg = ns["get_pos_and_neg_masks"](pl)
        sel = ns["OHNM_batch"](14, sc, pos, neg)
    save("model_loss_b14", pix_logits=inp["pix_logits"], link_logits=inp["link_logits"],
         pix_lab=inp["pix_lab"], link_lab=inp["link_lab"], loss=out.detach().numpy(),
         grad_pixel=yp.grad.numpy(), grad_link=yl.grad.numpy(),
         ohem_mask=sel.numpy().reshape(B, H, W))

    # NaN case (quirk Q2): a shard with no positive pixel at all.
    imgs = [synth.make_image(92, i, 8, 8, "G", "no_pos") for i in range(B)]
    inp = {k: np.stack([im[k] for im in imgs]) for k in imgs[0]}
    yp, yl = T(inp["pix_logits"], True), T(inp["link_logits"], True)
    out = ns["loss"](T(inp["pix_lab"]), yp, T(inp["link_lab"]), yl, torch.ones(B, 8, 8, 1))
    out.backward()
    # pixel logits are unconnected here (tf.cond takes no_pos -> constant 0): gradient None == zeros
    gp = yp.grad.numpy() if yp.grad is not None else np.zeros_like(inp["pix_logits"])
    save("model_loss_b14_nopos", pix_logits=inp["pix_logits"], link_logits=inp["link_logits"],
         pix_lab=inp["pix_lab"], link_lab=inp["link_lab"], loss=out.detach().numpy(),
         grad_pixel=gp, grad_link=yl.grad.numpy())

    # dice_coefficient (model.py:145-159)
    rng = np.random.default_rng(7)
    t = (rng.uniform(size=(3, 9, 11, 1)) > 0.7).astype(np.float32)
    p = rng.uniform(size=(3, 9, 11, 1)).astype(np.float32)
    m = (rng.uniform(size=(3, 9, 11, 1)) > 0.1).astype(np.float32)
    pt = T(p, True)
    d = ns["dice_coefficient"](T(t), pt, T(m))
    d.backward()
    save("dice_coefficient", t=t, p=p, m=m, loss=d.detach().numpy(), grad=pt.grad.numpy())


# ----------------------------------------------------------------------------- nets/model_vgg_16.py
def golden_vgg16():
    ns = ns_tf()
    for fn in ("dice_coefficient", "loss", "cal_link_loss", "ohem_loss"):
        src, span = cut("nets/model_vgg_16.py", fn)
        print("nets/model_vgg_16.py", fn, span)
        exec(src, ns)
    B, H, W = 4, 12, 20
    inp = synth.make_batch(93, B, H, W, "G")
    rng = np.random.default_rng(11)
    # dice head: predictions are probabilities (sigmoid outputs)
    pp = (1 / (1 + np.exp(-(inp["pix_logits"][..., 1:2] - inp["pix_logits"][..., 0:1])))).astype(np.float32)
    lk = inp["link_logits"].reshape(B, H, W, 8, 2)
    lp = (1 / (1 + np.exp(-(lk[..., 1] - lk[..., 0])))).astype(np.float32)
    m = (rng.uniform(size=(B, H, W, 1)) > 0.1).astype(np.float32)
    ppt, lpt = T(pp, True), T(lp, True)
    out = ns["loss"](T(inp["pix_lab"]), ppt, T(inp["link_lab"]), lpt, T(m))
    out.backward()
    save("vgg16_dice_loss", pix_lab=inp["pix_lab"], link_lab=inp["link_lab"], pix_prob=pp, link_prob=lp,
         training_mask=m, loss=out.detach().numpy(), grad_pixel=ppt.grad.numpy(), grad_link=lpt.grad.numpy())

    yp, yl = T(inp["pix_logits"], True), T(inp["link_logits"], True)
    out = ns["ohem_loss"](T(inp["pix_lab"]), yp, T(inp["link_lab"]), yl, T(m))
    out.backward()
    save("vgg16_ohem_loss", pix_logits=inp["pix_logits"], link_logits=inp["link_logits"],
         pix_lab=inp["pix_lab"], link_lab=inp["link_lab"], loss=out.detach().numpy(),
         grad_pixel=yp.grad.numpy(), grad_link=yl.grad.numpy())


# ----------------------------------------------------------------------------- nets/pixellink.py
def golden_pixellink_build_loss():
    src, span = cut("nets/pixellink.py", "build_loss", indent="    ")
    print("nets/pixellink.py build_loss", span)
    B, H, W = 5, 12, 20
    kinds = ["normal", "normal", "normal", "no_pos", "many_pos"]
    imgs = [synth.make_image(94, i, H, W, "G", kinds[i]) for i in range(B)]
    inp = {k: np.stack([im[k] for im in imgs]) for k in imgs[0]}
    config = types.SimpleNamespace(batch_size_per_gpu=B, max_neg_pos_ratio=3)
    ns = ns_tf()
    ns["config"] = config
    exec(src, ns)
    self = types.SimpleNamespace()
    self.pixel_cls = T(inp["pix_logits"], True)
    self.link_cls = T(inp["link_logits"], True)
    self.pixel_scores = tf.slim.softmax(self.pixel_cls)
    tf.reset_collections()
    # capture the diagnostic OHNM mask: reduce_sum(seg_selected_mask) is its only consumer (:155)
    captured = {}
    real_sum = tf.reduce_sum

    def spy_sum(x, axis=None):
        if isinstance(x, torch.Tensor) and x.dtype == torch.float32 and tuple(x.shape) == (B, H, W) \
                and "mask" not in captured and float(x.max()) <= 1.0 and float(x.min()) >= 0.0 \
                and bool(((x == 0) | (x == 1)).all()):
            captured["mask"] = x.detach().numpy().copy()
        return real_sum(x, axis)

    tf.reduce_sum = spy_sum
    try:
        ns["build_loss"](self, T(inp["pix_lab"][..., 0]), T(inp["link_lab"]))
    finally:
        tf.reduce_sum = real_sum
    losses = tf.get_collection(tf.GraphKeys.LOSSES)
    assert len(losses) == 2
    total = losses[0] + losses[1]
    total.backward()
    save("pixellink_build_loss", pix_logits=inp["pix_logits"], link_logits=inp["link_logits"],
         pix_lab=inp["pix_lab"], link_lab=inp["link_lab"],
         losses=np.array([l.detach().numpy() for l in losses]),
         grad_pixel=self.pixel_cls.grad.numpy(), grad_link=self.link_cls.grad.numpy(),
         ohem_mask=captured["mask"])


# ----------------------------------------------------------------------------- pure numpy pieces
def golden_numpy_pieces():
    ns = dict(np=np)
    src, span = cut("tool/pixellink_fn.py", "pixel_detect")
    print("tool/pixellink_fn.py pixel_detect", span)
    exec(src, ns)
    rng = np.random.default_rng(3)
    H, W = 24, 40
    score = rng.uniform(size=(1, H, W, 1)).astype(np.float32)
    score[0, 5:15, 5:30, 0] = rng.uniform(0.7, 1.0, (10, 25))
    geo = rng.uniform(0.6, 1.0, size=(8, 1, H, W, 2)).astype(np.float32)
    res = ns["pixel_detect"](score, geo)
    res2 = ns["pixel_detect"](score, geo, 0.75, 0.7)
    save("pixel_detect", score=score, geo=geo, res=res, res_075_07=res2)

    ns = dict(np=np)
    for fn in ("restore_rectangle_rbox", "restore_rectangle"):
        src, span = cut("datasets/icdar.py", fn)
        print("datasets/icdar.py", fn, span)
        exec(src, ns)
    N = 64
    origin = rng.uniform(0, 512, (N, 2)).astype(np.float32)
    geom = np.concatenate([rng.uniform(1, 80, (N, 4)), rng.uniform(-0.7, 0.7, (N, 1))], 1).astype(np.float32)
    geom[3, 4] = 0.0
    out = ns["restore_rectangle"](origin, geom)
    out_pos = ns["restore_rectangle"](origin[geom[:, 4] >= 0], geom[geom[:, 4] >= 0])
    out_empty = ns["restore_rectangle"](origin[:0], geom[:0])
    save("restore_rectangle", origin=origin, geometry=geom, out=out, out_pos_only=out_pos,
         out_empty_shape=np.array(out_empty.shape))

    import scipy.spatial.distance as dist
    ns = dict(np=np, dist=dist)
    for fn in ("order_points", "sort_poly"):
        src, span = cut("test.py", fn)
        print("test.py", fn, span)
        exec(src, ns)
    boxes = []
    for _ in range(40):
        c = rng.uniform(50, 400, 2)
        a = rng.uniform(-np.pi / 2, np.pi / 2)
        w, h = rng.uniform(10, 120), rng.uniform(5, 40)
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        p = (np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2) @ R.T + c
        p = np.roll(p, int(rng.integers(0, 4)), 0)
        boxes.append(p.astype(np.int64))
    boxes = np.stack(boxes)
    save("order_points", boxes=boxes, ordered=np.stack([ns["order_points"](b) for b in boxes]),
         sorted_poly=np.stack([ns["sort_poly"](b) for b in boxes]))

    # example.py:5,12 known answer: softmax of [[1,2],[3,4],[5,6],[7,8]] rows
    x = torch.tensor([[[[1., 2.], [3., 4.], [5., 6.], [7., 8.]]]])
    save("example_softmax", x=x.numpy(), y=tf.slim.softmax(x).numpy())


def golden_generate_rbox():
    """tool/pixellink_fn.py valid_link + generate_rbox, executed as written except for the two Python-2
    idioms that do not run under Python 3: `h/4`, `w/4` (integer division there) and `zip(...)` (a list there)."""
    import cv2
    ns = dict(np=np, cv2=cv2)
    src, span = cut("tool/pixellink_fn.py", "valid_link")
    print("tool/pixellink_fn.py valid_link", span)
    exec(src, ns)
    src, span = cut("tool/pixellink_fn.py", "generate_rbox")
    print("tool/pixellink_fn.py generate_rbox", span)
    assert "new_h = h/4" in src and "new_w = w/4" in src and "points = zip(" in src
    src = src.replace("new_h = h/4", "new_h = h//4").replace("new_w = w/4", "new_w = w//4")
    src = src.replace("points = zip(xs[idx, :] * w, ys[idx, :] * h )", "points = list(zip(xs[idx, :] * w, ys[idx, :] * h ))")
    exec(src, ns)
    rng = np.random.default_rng(11)
    cases = {}
    for ci, (h, w, n) in enumerate([(64, 96, 3), (128, 128, 6), (48, 200, 9)]):
        xs, ys = [], []
        for i in range(n):
            c = rng.uniform(0.1, 0.9, 2)
            a = rng.uniform(-0.6, 0.6)
            hw, hh = rng.uniform(0.05, 0.3), rng.uniform(0.03, 0.12)
            R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
            p = (np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T) + c   # may leave [0,1]: clipped by fillPoly
            xs.append(p[:, 0]); ys.append(p[:, 1])
        xs, ys = np.array(xs, np.float32), np.array(ys, np.float32)
        bboxes = np.stack([ys.min(1), xs.min(1), ys.max(1), xs.max(1)], 1).astype(np.float32)
        ignored = (rng.uniform(size=n) < 0.3).astype(np.int64)
        score, link, show = ns["generate_rbox"](h, w, xs, ys, bboxes, ignored)
        cases.update({"h%d" % ci: h, "w%d" % ci: w, "xs%d" % ci: xs, "ys%d" % ci: ys, "bboxes%d" % ci: bboxes,
                      "ignored%d" % ci: ignored, "score%d" % ci: score, "link%d" % ci: link, "show%d" % ci: show})
    save("generate_rbox", n_cases=3, **cases)


# ----------------------------------------------------------------------------- D2/D3: link graph + DFS + boxes
def cut_lines(path, lo, hi):
    """Lines lo..hi (1-based, inclusive) of a reference file, dedented."""
    lines = open(os.path.join(REF, path)).read().split("\n")
    return textwrap.dedent("\n".join(lines[lo - 1:hi]))


def _py3(src):
    """The only edits made to the inline decode blocks: Python-2 `print x` statements dropped, dict.has_key ->
    `in`, and np.int0 (removed in numpy 2) -> np.intp.  Returns (patched source, number of edits)."""
    n = 0
    out = []
    for l in src.split("\n"):
        if re.match(r"^\s*print\s+[^(]", l):
            l = l[: len(l) - len(l.lstrip())] + "pass"
            n += 1
        if "graph.has_key(v)" in l:
            l = l.replace("graph.has_key(v)", "(v in graph)")
            n += 1
        if "np.int0(" in l:
            l = l.replace("np.int0(", "np.intp(")
            n += 1
        out.append(l)
    return "\n".join(out), n


def _symmetric_flag_maps(seed, H, W, rects, min_side=3):
    """P [H,W] bool, L [H,W,8] bool with symmetric link decisions (L[v,d] == L[u,opp(d)]) and no positive on
    the map border: on such maps the reference's directed, seed-order dependent DFS is order independent
    (quirk Q10), so its Python-3 execution is THE reference result."""
    import cv2
    rng = np.random.default_rng(seed)
    P = np.zeros((H, W), np.uint8)
    keep = np.zeros((H, W), np.float32)          # per-pixel edge survival probability
    for (cx, cy, lw, lh, ang, p_edge) in rects:
        c, s = np.cos(ang), np.sin(ang)
        pts = (np.array([[-lw, -lh], [lw, -lh], [lw, lh], [-lw, lh]]) / 2.0) @ np.array([[c, -s], [s, c]]).T + (cx, cy)
        m = np.zeros((H, W), np.uint8)
        cv2.fillPoly(m, [np.round(pts).astype(np.int32)], 1)
        P |= m
        keep[m > 0] = p_edge
    P[0, :] = P[-1, :] = 0
    P[:, 0] = P[:, -1] = 0
    P = P.astype(bool)
    P &= rng.uniform(size=(H, W)) > 0.03         # pinholes
    L = np.zeros((H, W, 8), bool)
    for d in (3, 4, 5, 7):                       # each undirected pair once
        dy, dx = synth.NEIGHBOURS[d]
        o = synth.OPPOSITE[d]
        on = rng.uniform(size=(H, W)) < keep
        ys, xs = np.nonzero(on)
        ok = (ys + dy >= 0) & (ys + dy < H) & (xs + dx >= 0) & (xs + dx < W)
        ys, xs = ys[ok], xs[ok]
        L[ys, xs, d] = True
        L[ys + dy, xs + dx, o] = True
    # link decisions of background pixels are arbitrary (the graph never reads them): noise
    # (kept to a 2-pixel band around the text so that the packed fixture stays small)
    band = cv2.dilate(P.astype(np.uint8), np.ones((5, 5), np.uint8)).astype(bool) & ~P
    L |= band[:, :, None] & (rng.uniform(size=(H, W, 8)) < 0.3)
    return P, L


def _run_inline_decode(path, spans, H, W, P, L, hi_score, lo_score, flags):
    """exec the inline decode block of a reference script on score maps built from (P, L)."""
    import cv2
    src = "\n".join(cut_lines(path, lo, hi) for lo, hi in spans)
    src, nedit = _py3(src)
    assert ("%d" % W) in src and ("%d" % H) in src, "map-size literals of the script"
    pixel_score = np.where(P, hi_score, lo_score).astype(np.float32)
    ns = dict(np=np, cv2=cv2, FLAGS=flags, pixel_score=pixel_score,
              link_score_set=[np.where(L[:, :, d], hi_score, lo_score).astype(np.float32) for d in range(8)],
              im_ori=np.zeros((8, 8, 3), np.uint8), boxes=[])
    exec(src, ns)
    assert np.array_equal(ns["pixel_seg"], P)
    return ns, nedit


def golden_link_graph():
    """test_pixellink_fast.py:111,115-178,191-202 (192x320 maps, > 10 px, scale 4.0 / 3.75) and
    test_pixellink.py:113,117-181,206-217 (720x1280, > 200 px, unscaled), executed by line range."""
    flags = types.SimpleNamespace(pixel_conf_threshold=0.8, link_conf_threshold=0.9)
    out = {}
    # ---- 4s script
    H, W = 192, 320
    rng = np.random.default_rng(5)
    rects = []
    for i in range(14):
        rects.append((rng.uniform(20, W - 20), rng.uniform(15, H - 15), rng.uniform(8, 70), rng.uniform(2, 12),
                      rng.uniform(-0.8, 0.8), [1.0, 0.7, 0.45, 0.3][i % 4]))
    P, L = _symmetric_flag_maps(41, H, W, rects)
    ns, nedit = _run_inline_decode("test_pixellink_fast.py", [(111, 111), (115, 178), (191, 202)], H, W, P, L,
                                   0.95, 0.05, flags)
    print("test_pixellink_fast.py inline decode: %d groups, %d py3 edits" % (ns["gid"] - 1, nedit))
    out.update(_pack_case("fast", H, W, P, L, ns, 10, (1280.0 / 320, 720.0 / 192)))
    # the ICDAR result-file format, test_pixellink_fast.py:209-217, written by the script's own lines
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        ns["FLAGS"].output_dir = td
        ns["image_name"] = "/some/dir/img_7.jpg"
        ns["os"] = os
        exec(_py3(cut_lines("test_pixellink_fast.py", 209, 217))[0], ns)
        out["fast_res_txt"] = np.frombuffer(open(os.path.join(td, "res_img_7.txt"), "rb").read(), np.uint8)
    # ---- full-resolution script (small components only: its DFS is quadratic in the component size)
    H, W = 720, 1280
    rng = np.random.default_rng(6)
    rects = []
    for i in range(12):
        rects.append((rng.uniform(60, W - 60), rng.uniform(40, H - 40), rng.uniform(20, 110), rng.uniform(4, 22),
                      rng.uniform(-0.8, 0.8), [1.0, 0.6, 0.4][i % 3]))
    P, L = _symmetric_flag_maps(42, H, W, rects)
    ns, nedit = _run_inline_decode("test_pixellink.py", [(113, 113), (117, 181), (206, 217)], H, W, P, L,
                                   0.95 * 255, 0.05 * 255, flags)
    print("test_pixellink.py inline decode: %d groups, %d py3 edits" % (ns["gid"] - 1, nedit))
    out.update(_pack_case("full", H, W, P, L, ns, 200, (1.0, 1.0)))
    save("link_graph", **out)


def _pack_case(tag, H, W, P, L, ns, min_size, scale):
    """Canonicalise the script's result: gid g -> the group's minimum pixel index; boxes in ascending-label order."""
    group = np.asarray(ns["group_idx"]).reshape(H, W)
    gid = int(ns["gid"])
    labels = np.full((H, W), -1, np.int32)
    mins = []
    for g in range(1, gid):
        idx = np.flatnonzero(group.reshape(-1) == g)
        mins.append(int(idx.min()))
        labels.reshape(-1)[idx] = idx.min()
    order = np.argsort(mins)
    boxes = np.stack([np.asarray(ns["boxes"][i]) for i in order]).astype(np.int64) if mins else np.zeros((0, 4, 2), np.int64)
    return {tag + "_H": H, tag + "_W": W, tag + "_P": np.packbits(P), tag + "_L": np.packbits(L),
            tag + "_labels": labels, tag + "_boxes": boxes, tag + "_min_size": min_size,
            tag + "_scale": np.asarray(scale, np.float64), tag + "_n_groups": gid - 1}


def golden_icdar_generate_rbox():
    """datasets/icdar.py valid_link (:83-105) and generate_rbox (:486-539) executed as written (FLAGS.min_text_size
    = 10, :25), on overlapping polygons (the loop is order dependent) that touch all four borders (index -1 wraps;
    x == h-1 / y == w-1 return early), plus the [::4, ::4] subsample of :632-634."""
    import cv2
    ns = dict(np=np, cv2=cv2, FLAGS=types.SimpleNamespace(min_text_size=10))
    for fn in ("valid_link", "generate_rbox"):
        src, span = cut("datasets/icdar.py", fn)
        print("datasets/icdar.py", fn, span)
        exec(src, ns)
    rng = np.random.default_rng(21)
    out = {}
    for ci, (size, n) in enumerate([(64, 5), (96, 9), (128, 12)]):
        polys = []
        for i in range(n):
            c = rng.uniform(-0.05, 1.05, 2) * size
            a = rng.uniform(-0.7, 0.7)
            hw, hh = rng.uniform(0.05, 0.35) * size, rng.uniform(0.02, 0.15) * size
            R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
            polys.append((np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T + c).astype(np.float32))
        polys = np.stack(polys)
        tags = rng.uniform(size=n) < 0.3
        score, geo, tm = ns["generate_rbox"]((size, size), polys, tags)
        out.update({"size%d" % ci: size, "polys%d" % ci: polys, "tags%d" % ci: tags, "score%d" % ci: score,
                    "geo%d" % ci: geo, "tmask%d" % ci: tm, "geo4s%d" % ci: geo[::4, ::4, :].astype(np.float32),
                    "score4s%d" % ci: score[::4, ::4].astype(np.float32), "tmask4s%d" % ci: tm[::4, ::4].astype(np.float32)})
    save("icdar_generate_rbox", n_cases=3, **out)


# ----------------------------------------------------------------------------- tool/bboxes.py + tool/metrics.py (N4)
class _NpTensor(np.ndarray):
    def set_shape(self, _shape):
        pass


class _NpTensorArray:
    def __init__(self, dtype, size, dynamic_size=False, infer_shape=True):
        self.items = [None] * int(size)

    def write(self, i, v):
        self.items[int(i)] = np.asarray(v)
        return self

    def stack(self):
        return np.stack(self.items) if self.items else np.zeros((0,), bool)


def _np_tf():
    """Eager numpy stand-in for the handful of TF ops tool/bboxes.py::bboxes_matching and tool/metrics.py use."""
    import contextlib

    t = types.SimpleNamespace()
    t.bool, t.int32, t.int64, t.float32 = np.bool_, np.int32, np.int64, np.float32
    t.name_scope = lambda *a, **k: contextlib.nullcontext()
    t.cast = lambda x, dtype=None: np.asarray(x).astype(dtype)
    t.count_nonzero = lambda x: np.int64(np.count_nonzero(x))
    t.logical_not, t.logical_and, t.logical_or = np.logical_not, np.logical_and, np.logical_or
    t.zeros = lambda shape, dtype=np.float32: np.zeros(tuple(int(v) for v in np.atleast_1d(shape)), dtype)
    t.zeros_like = np.zeros_like
    t.shape = lambda x: np.asarray(np.shape(x), np.int32)
    t.size = lambda x: np.int32(np.size(x))
    t.range = lambda n, dtype=np.int32: np.arange(int(n), dtype=dtype)
    t.TensorArray = _NpTensorArray
    t.less = np.less
    t.greater = np.greater
    t.equal = np.equal
    t.divide = lambda a, b: np.divide(a, b, out=np.zeros_like(np.asarray(a, np.float32)), where=np.asarray(b) != 0)
    t.argmax = lambda x, axis=0: np.int64(np.argmax(x, axis=axis))
    t.reshape = lambda x, shp: np.reshape(x, tuple(int(v) for v in shp))
    t.reduce_sum = lambda x, axis=None: np.sum(x, axis=axis, dtype=np.asarray(x).dtype)
    t.where = lambda c, a, b, name=None: np.where(c, a, b)
    t.tuple = lambda xs: tuple(xs)

    def py_func(fn, inputs, _tout):
        return np.asarray(fn(*[np.asarray(v) for v in inputs])).view(_NpTensor)

    def while_loop(cond, body, loop_vars, parallel_iterations=1, back_prop=False):
        vs = list(loop_vars)
        while cond(*vs):
            vs = list(body(*vs))
        return vs

    t.py_func, t.while_loop = py_func, while_loop
    return t


def golden_evaluation():
    """tool/bboxes.py np_bboxes_jaccard (:252-282), bboxes_jaccard (:247-250), bboxes_matching (:158-246) and
    tool/metrics.py precision_recall (:66-80), fmean (:82-85) with tool/math.py safe_divide (:27-41), executed as
    written.  `util.img` (un-vendored dengdan/pylib) is its three one-line cv2/numpy wrappers."""
    import cv2

    util = types.SimpleNamespace(img=types.SimpleNamespace(
        points_to_contours=lambda pts: [np.asarray(pts, np.int32).reshape(-1, 1, 2)],
        black=lambda shape: np.zeros(tuple(int(v) for v in shape), np.uint8),
        draw_contours=lambda img, contours, idx=-1, color=1, border_width=1: cv2.drawContours(img, contours, idx, color, border_width)))
    tfn = _np_tf()
    ns = dict(np=np, cv2=cv2, util=util, tf=tfn, zip=lambda *a: list(zip(*a)))
    for fn in ("np_bboxes_jaccard", "bboxes_jaccard", "bboxes_matching"):
        src, span = cut("tool/bboxes.py", fn)
        print("tool/bboxes.py", fn, span)
        exec(src, ns)
    mns = dict(tf=tfn, math_ops=tfn)
    src, span = cut("tool/math.py", "safe_divide")
    print("tool/math.py safe_divide", span)
    exec(src, mns)
    mns["tfe_math"] = types.SimpleNamespace(safe_divide=mns["safe_divide"])
    for fn in ("precision_recall", "fmean"):
        src, span = cut("tool/metrics.py", fn)
        print("tool/metrics.py", fn, span)
        exec(src, mns)

    rng = np.random.default_rng(44)

    def quad(c, w, h, a):
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        return np.maximum((np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2 @ R.T + c), 0).astype(np.int32)

    out = {}
    n_cases = 6
    for ci in range(n_cases):
        G = int(rng.integers(1, 14))
        size = (160, 90) if ci % 2 else (1280, 720)
        gts = []
        for g in range(G):
            c = rng.uniform(0.1, 0.9, 2) * size
            gts.append(quad(c, rng.uniform(0.03, 0.3) * size[0], rng.uniform(0.02, 0.12) * size[1], rng.uniform(-0.6, 0.6)))
        gts = np.stack(gts)
        dets = []
        for g in range(G):          # detections: jittered copies (some twice: the second one is a false positive), plus strays
            for rep in range(int(rng.integers(0, 3))):
                j = rng.normal(0, 0.006 * size[0] * (1 + 5 * (rng.uniform() < 0.3)), (4, 2))
                dets.append(np.maximum(gts[g] + j, 0).astype(np.int32))
        for k in range(int(rng.integers(1, 5))):
            dets.append(quad(rng.uniform(0.1, 0.9, 2) * size, rng.uniform(0.03, 0.2) * size[0],
                             rng.uniform(0.02, 0.1) * size[1], rng.uniform(-1.5, 1.5)))
        if ci == 0:                  # degenerate shapes: a segment, a point, a bow-tie
            dets += [np.array([[5, 5], [40, 9], [40, 9], [5, 5]], np.int32), np.array([[7, 7]] * 4, np.int32),
                     np.array([[10, 10], [60, 40], [60, 10], [10, 40]], np.int32)]
        order = rng.permutation(len(dets))
        dets = np.stack([dets[i] for i in order])
        gign = (rng.uniform(size=G) < 0.25).astype(np.int32)
        gxs, gys = gts[:, :, 0].copy(), gts[:, :, 1].copy()
        bboxes = dets.reshape(-1, 8)
        jac = np.stack([ns["np_bboxes_jaccard"](bboxes[i], gxs, gys) for i in range(len(bboxes))])
        n_g, tp, fp = ns["bboxes_matching"](bboxes, gxs, gys, gign, matching_threshold=0.5)
        pre, rec = mns["precision_recall"](n_g, tp, fp)
        with np.errstate(invalid="ignore", divide="ignore"):
            fm = mns["fmean"](pre, rec)
        out.update({"bboxes%d" % ci: bboxes, "gxs%d" % ci: gxs, "gys%d" % ci: gys, "gignored%d" % ci: gign,
                    "jaccard%d" % ci: jac.astype(np.float32), "n_gbboxes%d" % ci: np.int64(n_g), "tp%d" % ci: tp,
                    "fp%d" % ci: fp, "precision%d" % ci: np.float32(pre), "recall%d" % ci: np.float32(rec),
                    "fmean%d" % ci: np.float32(fm)})
        print("case", ci, "G", G, "D", len(bboxes), "tp", int(tp.sum()), "fp", int(fp.sum()), "n_g", int(n_g), "P/R/F", pre, rec, fm)
    save("evaluation", n_cases=n_cases, **out)


# ----------------------------------------------------------------------------- the head's logit producer (N3)
class _NpSlim:
    """Eager float64 stand-in for slim.conv2d (1x1 only) / arg_scope / batch_norm, weights supplied by the test:
    by `scope` name when the call names one (nets/pixellink.py), else in call order (nets/model.py)."""

    def __init__(self, by_scope=None, in_order=None):
        self.by_scope, self.in_order, self.calls, self.defaults = by_scope, in_order, [], [{}]
        self.batch_norm = "batch_norm"

    def l2_regularizer(self, _wd):
        return None

    def arg_scope(self, _fns, **kw):
        import contextlib

        @contextlib.contextmanager
        def cm():
            self.defaults.append({**self.defaults[-1], **kw})
            try:
                yield
            finally:
                self.defaults.pop()
        return cm()

    def conv2d(self, x, num_outputs, kernel_size, scope=None, **kw):
        args = {**self.defaults[-1], **kw}
        assert kernel_size in (1, [1, 1]), kernel_size
        rec = self.by_scope[scope] if scope is not None else self.in_order[len(self.calls)]
        self.calls.append(scope)
        w = np.asarray(rec["w"], np.float64)
        assert w.shape == (x.shape[-1], num_outputs), (w.shape, x.shape, num_outputs)
        y = np.asarray(x, np.float64) @ w
        if args.get("normalizer_fn") is not None:      # inference-mode batch norm folded to scale / shift
            y = y * np.asarray(rec["scale"], np.float64) + np.asarray(rec["shift"], np.float64)
        else:
            y = y + np.asarray(rec["b"], np.float64)
        act = args.get("activation_fn")
        return act(y) if act is not None else y


def _np_tf_image():
    from oracle.head_logits import resize_bilinear_x2
    t = types.SimpleNamespace()
    t.shape = lambda x: np.asarray(np.shape(x))

    def resize_bilinear(x, size):
        assert tuple(int(v) for v in size) == (2 * x.shape[1], 2 * x.shape[2]), size
        return resize_bilinear_x2(x)     # TF op semantics restated (oracle/head_logits.py)
    t.image = types.SimpleNamespace(resize_bilinear=resize_bilinear)
    t.nn = types.SimpleNamespace(relu=lambda v: np.maximum(v, 0.0))
    t.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=lambda: None))
    t.zeros_initializer = lambda: None
    return t


def golden_head_logits():
    """nets/pixellink.py:37-38 unpool + :57-67 (the body of _add_pixellink_layers up to link_cls) and
    nets/model.py:14-15 unpool + :129-141 (pixel_1 .. link_4 inside the arg_scope of :103-107), executed as written
    over eager float64 stand-ins for slim.conv2d / tf.image.resize_bilinear."""
    rng = np.random.default_rng(77)
    tfi = _np_tf_image()
    out = {}
    # ---- nets/pixellink.py
    B, H, W = 1, 16, 24           # 1/4-resolution map; conv4_3 at 1/8, conv5_3 / fc7 at 1/16
    chans = {"fc7": 40, "conv5_3": 32, "conv4_3": 24, "conv3_3": 20}
    ep = {"fc7": rng.standard_normal((B, H // 4, W // 4, chans["fc7"])), "conv5_3": rng.standard_normal((B, H // 4, W // 4, chans["conv5_3"])),
          "conv4_3": rng.standard_normal((B, H // 2, W // 2, chans["conv4_3"])), "conv3_3": rng.standard_normal((B, H, W, chans["conv3_3"]))}
    ep = {k: v.astype(np.float32) for k, v in ep.items()}
    weights = {}
    for kind, n, last in (("pixel", 2, "text_predication"), ("link", 16, "link_predication")):
        for st, name in ((6, "fc7"), (5, "conv5_3"), (4, "conv4_3"), (3, "conv3_3")):
            weights["stage_%d_%s_fuse" % (st, kind)] = dict(w=(rng.standard_normal((chans[name], n)) / np.sqrt(chans[name])).astype(np.float32),
                                                             b=(0.1 * rng.standard_normal(n)).astype(np.float32))
        weights[last] = dict(w=(rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32), b=(0.1 * rng.standard_normal(n)).astype(np.float32))
    slim_ = _NpSlim(by_scope=weights)
    ns = dict(tf=tfi, slim=slim_)
    exec(cut_lines("nets/pixellink.py", 37, 38), ns)                       # def unpool(self, inputs)
    self_ = types.SimpleNamespace(weight_decay=1e-5, unpool=lambda x: ns["unpool"](None, x))
    body = cut_lines("nets/pixellink.py", 57, 67)
    print("nets/pixellink.py:57-67\n" + "\n".join("   | " + l[:110] for l in body.split("\n")[:3]) + "\n   | ...")
    ns.update(self=self_, end_points=ep)
    exec(body, ns)
    assert len(slim_.calls) == 10, slim_.calls
    for k, v in ep.items():
        out["pl_" + k] = v
    for k, v in weights.items():
        out["pl_w_" + k], out["pl_b_" + k] = v["w"], v["b"]
    out["pl_pixel_cls"], out["pl_link_cls"] = ns["pixel_cls"].astype(np.float32), ns["link_cls"].astype(np.float32)
    # ---- nets/model.py
    B, H, W = 1, 16, 16
    fch = [40, 32, 24, 12]        # pool5 (1/32), pool4, pool3, pool2 (1/4)
    fm = [rng.standard_normal((B, H // 8, W // 8, fch[0])), rng.standard_normal((B, H // 4, W // 4, fch[1])),
          rng.standard_normal((B, H // 2, W // 2, fch[2])), rng.standard_normal((B, H, W, fch[3]))]
    fm = [v.astype(np.float32) for v in fm]
    order = []
    recs = {}
    for kind, n in (("pixel", 2), ("link", 16)):     # call order of :129-141: pixel f0..f3, link f0..f3, pixel_4, link_4
        for i in range(4):
            r = dict(w=(rng.standard_normal((fch[i], n)) / np.sqrt(fch[i])).astype(np.float32),
                     scale=rng.uniform(0.5, 1.5, n).astype(np.float32), shift=(0.3 * rng.standard_normal(n)).astype(np.float32))
            recs["%s_f%d" % (kind, i)] = r
            order.append(r)
    for kind, n in (("pixel", 2), ("link", 16)):
        r = dict(w=(rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32), b=(0.1 * rng.standard_normal(n)).astype(np.float32))
        recs["%s_out" % kind] = r
        order.append(r)
    slim_ = _NpSlim(in_order=order)
    ns = dict(tf=tfi, slim=slim_, np=np, feature_maps=fm, PIXEL_OUTPUT=2, LINK_OUTPUT=16, print=lambda *a, **k: None)
    exec(cut_lines("nets/model.py", 14, 15), ns)                            # def unpool(inputs)
    body = cut_lines("nets/model.py", 129, 141)
    print("nets/model.py:129-141\n" + "\n".join("   | " + l[:110] for l in body.split("\n")[:2]) + "\n   | ...")
    with slim_.arg_scope([slim_.conv2d], activation_fn=tfi.nn.relu, normalizer_fn=slim_.batch_norm, normalizer_params={}):   # :103-107
        exec(body, ns)
    assert len(slim_.calls) == 10
    for i, v in enumerate(fm):
        out["md_f%d" % i] = v
    for k, r in recs.items():
        for kk, vv in r.items():
            out["md_%s_%s" % (k, kk)] = vv
    out["md_pixel_4"], out["md_link_4"] = ns["pixel_4"].astype(np.float32), ns["link_4"].astype(np.float32)
    save("head_logits", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only == "rbox":
        golden_generate_rbox()
    elif only == "link_graph":
        golden_link_graph()
    elif only == "icdar":
        golden_icdar_generate_rbox()
    elif only == "evaluation":
        golden_evaluation()
    elif only == "head_logits":
        golden_head_logits()
    else:
        golden_model_loss()
        golden_vgg16()
        golden_pixellink_build_loss()
        golden_numpy_pieces()
        golden_generate_rbox()
        golden_link_graph()
        golden_icdar_generate_rbox()
        golden_evaluation()
        golden_head_logits()
