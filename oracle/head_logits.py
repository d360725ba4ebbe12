"""CPU oracle for the head's logit producer — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package; the product
(tensorflow_ocr_b200/) never does.

Restates, in float64 numpy,
  * nets/pixellink.py:37-38 unpool and :56-67 _add_pixellink_layers — the PixelLink-4s fusion:
        s1 = conv1x1(fc7) + conv1x1(conv5_3);  s2 = unpool(s1) + conv1x1(conv4_3);
        s3 = unpool(s2) + conv1x1(conv3_3);    logits = conv1x1(s3)
    for the 2 pixel and the 16 link channels (plain 1x1 convolutions with bias, no activation);
  * nets/model.py:14-15 unpool and :129-141 — the EAST-fork fusion over pool5..pool2, whose fuse convolutions
    carry slim's arg_scope (batch norm + ReLU, :103-107) and whose last convolution is plain; batch norm in its
    INFERENCE form (moving statistics), i.e. a per-channel scale and shift after the convolution.

TensorFlow is the un-vendored dependency (README.md:2 "tensorflow 1.4"); its op semantics are restated from the
documented behaviour: slim.conv2d 1x1 = per-pixel matrix product (+ bias | batch norm) (+ activation);
tf.image.resize_bilinear(align_corners=False): src = dst * (in / out), lower index floor(src), upper index
min(lower + 1, in - 1), linear weights.  The STRUCTURE (which feature feeds which stage, where the unpools sit)
is pinned by executing the reference's own lines over these op restatements
(tests/golden/make_golden.py::golden_head_logits -> tests/golden/head_logits.npz).
"""
from __future__ import annotations

import numpy as np


def resize_bilinear_x2(x):
    """tf.image.resize_bilinear(x, [2H, 2W]) (align_corners=False), NHWC."""
    x = np.asarray(x, np.float64)
    B, H, W, C = x.shape

    def axis(n_in):
        src = np.arange(2 * n_in, dtype=np.float64) * (n_in / (2.0 * n_in))
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        return lo, hi, src - lo

    ylo, yhi, fy = axis(H)
    xlo, xhi, fx = axis(W)
    top = x[:, ylo][:, :, xlo] * (1 - fx)[None, None, :, None] + x[:, ylo][:, :, xhi] * fx[None, None, :, None]
    bot = x[:, yhi][:, :, xlo] * (1 - fx)[None, None, :, None] + x[:, yhi][:, :, xhi] * fx[None, None, :, None]
    return top * (1 - fy)[None, :, None, None] + bot * fy[None, :, None, None]


def conv1x1(x, w, scale=None, shift=None, relu=False):
    """slim.conv2d(x, n, 1): x [B,H,W,K], w [K,n]; then `* scale + shift` (bias: scale None; folded batch norm), ReLU."""
    y = np.asarray(x, np.float64) @ np.asarray(w, np.float64)
    if scale is not None:
        y = y * np.asarray(scale, np.float64)
    if shift is not None:
        y = y + np.asarray(shift, np.float64)
    return np.maximum(y, 0.0) if relu else y


def pixellink_layers(end_points, p):
    """nets/pixellink.py:56-67.  end_points: fc7, conv5_3, conv4_3, conv3_3 (NHWC); p[scope] = (w [K,n], b [n]) for the
    ten scopes stage_{6,5,4,3}_{pixel,link}_fuse, text_predication, link_predication.  -> pixel_cls [B,H,W,2],
    link_cls [B,H,W,16] (float64)."""
    out = []
    for kind, last in (("pixel", "text_predication"), ("link", "link_predication")):
        c = lambda name, scope: conv1x1(end_points[name], p[scope][0], None, p[scope][1])
        s1 = c("fc7", "stage_6_%s_fuse" % kind) + c("conv5_3", "stage_5_%s_fuse" % kind)
        s2 = resize_bilinear_x2(s1) + c("conv4_3", "stage_4_%s_fuse" % kind)
        s3 = resize_bilinear_x2(s2) + c("conv3_3", "stage_3_%s_fuse" % kind)
        out.append(conv1x1(s3, p[last][0], None, p[last][1]))
    return out[0], out[1]


def model_head(feature_maps, p):
    """nets/model.py:129-141.  feature_maps = [pool5, pool4, pool3, pool2]; p["pixel"] / p["link"] = list of five
    (w, scale, shift) triples: the four fuse convolutions (batch norm folded to scale/shift, then ReLU) in the order
    f0..f3 and the plain last convolution (scale None, shift = bias).  -> pixel_4 [B,H,W,2], link_4 [B,H,W,16]."""
    out = []
    for kind in ("pixel", "link"):
        q = p[kind]
        f = lambda i: conv1x1(feature_maps[i], q[i][0], q[i][1], q[i][2], relu=True)
        s1 = resize_bilinear_x2(f(0)) + f(1)
        s2 = resize_bilinear_x2(s1) + f(2)
        s3 = resize_bilinear_x2(s2) + f(3)
        out.append(conv1x1(s3, q[4][0], q[4][1], q[4][2]))
    return out[0], out[1]
