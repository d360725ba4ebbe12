"""Oracle: PixelLink head losses and their gradients (numpy restatement).

TEST INFRASTRUCTURE — see oracle/__init__.py.  Follows, line by line:

* nets/model.py:145-261      (dice_coefficient, OHNM_single_image, OHNM_batch,
                              get_pos_and_neg_masks, loss)              L1-L5, L7
* nets/pixellink.py:88-263   (PixelLinkNet.build_loss)                  L3', L6
* nets/model_vgg_16.py:179-282 (dice loss, cal_link_loss, ohem_loss)    L8, L9
* focal loss (Lin et al. 2017) — NOT IN THE REFERENCE (README.md:3 names it
  only): restated from the paper, PARITY UNPINNED.                       L10

TensorFlow 1.4 op semantics restated (un-vendored dependency):
``slim.softmax``  = exp(x-max)/sum exp(x-max) in fp32;
``sparse_softmax_cross_entropy_with_logits`` = log(sum exp(x-max)) - (x_label-max),
gradient softmax - onehot; ``top_k(-v, k)[-1]`` = minus the k-th smallest of v;
masks / counts / thresholds carry no gradient.

Element-wise arithmetic is fp32 like the reference; reductions accumulate in
fp64 (the most accurate value any fp32 summation order can be compared to
within the 1e-5 relative tolerance the contract states).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------- TF op restatements
def softmax2(x):
    """slim.softmax over the last (size-2) axis, fp32 (model.py:216; pixellink.py:71)."""
    x = np.asarray(x, f32)
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m, dtype=f32)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=f32)).astype(f32)


def sparse_xent2(logits, labels):
    """tf.nn.sparse_softmax_cross_entropy_with_logits on [...,2] logits, int labels."""
    x = np.asarray(logits, f32)
    m = np.max(x, axis=-1, keepdims=True)
    sh = (x - m).astype(f32)
    lse = np.log(np.sum(np.exp(sh, dtype=f32), axis=-1, dtype=f32), dtype=f32)
    lab = np.asarray(labels).astype(np.int64)
    picked = np.take_along_axis(sh, lab[..., None], axis=-1)[..., 0]
    return (lse - picked).astype(f32)


def _sum(x):
    return float(np.sum(np.asarray(x), dtype=np.float64))


def _div(a, b):
    """fp32 a/b with IEEE semantics (0/0 = NaN, x/0 = inf) and no warnings."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return f32(f32(a) / f32(b))


# --------------------------------------------------------------------------- L1
def get_pos_and_neg_masks(labels):
    """model.py:199-202 — on int32-cast labels."""
    labels = np.asarray(labels)
    return labels == 1, labels == 0


# --------------------------------------------------------------------------- L3 / L3'
def kth_smallest(v, k):
    """k-th smallest (1-based) of a 1-D array == -top_k(-v, k)[-1] (model.py:176-177)."""
    v = np.asarray(v)
    return np.partition(v, k - 1)[k - 1]


def OHNM_single_image(scores, n_pos, neg_mask, ratio=3):
    """model.py:161-184.  ``scores`` = softmax prob of the NEGATIVE class.

    n_pos > 0: k = min(ratio*n_pos, #neg); thr = k-th smallest negative score;
    selected = neg & (scores <= thr) (ties at thr all selected).  k == 0 (no
    negatives) raises in TF (``vals[-1]`` of an empty tensor); defined here as
    "select none" (SURVEY §8a L3).  Returns (mask fp32, thr or nan).
    """
    scores = np.asarray(scores, f32)
    neg_mask = np.asarray(neg_mask, bool)
    if n_pos > 0:
        n_neg = min(int(n_pos) * ratio, int(neg_mask.sum()))
        if n_neg == 0:
            return np.zeros(neg_mask.shape, f32), f32(np.nan)
        thr = kth_smallest(scores[neg_mask], n_neg)
        return (neg_mask & (scores <= thr)).astype(f32), f32(thr)
    return np.zeros(neg_mask.shape, f32), f32(np.nan)


def OHNM_single_image_pixellink(scores, n_pos, neg_mask, max_neg_pos_ratio=3):
    """nets/pixellink.py:106-136 (the ``tf.where``-padded variant, quirk Q5).

    k = min(ratio*n_pos, max(#neg, 1)); scores of non-negatives are replaced by 0
    BEFORE top_k, so the zeros occupy the smallest slots.
    """
    scores = np.asarray(scores, f32).reshape(-1)
    neg = np.asarray(neg_mask, bool).reshape(-1)
    if n_pos > 0:
        n_neg = min(int(n_pos) * max_neg_pos_ratio, max(int(neg.sum()), 1))
        neg_conf = np.where(neg, scores, f32(0))
        thr = kth_smallest(neg_conf, n_neg)
        sel = (neg & (scores <= thr)).astype(f32)
        return sel.reshape(np.shape(neg_mask)), f32(thr)
    return np.zeros(np.shape(neg_mask), f32), f32(np.nan)


def OHNM_batch(batch_size, neg_conf, pos_mask, neg_mask, ratio=3, variant="model"):
    """model.py:186-197 / pixellink.py:138-150.

    The reference passes the literal 14 (model.py:220, quirk Q1); ``batch_size``
    is accepted and ignored in favour of ``pos_mask.shape[0]``.
    Returns (selected_mask fp32 [B, ...], thr [B]).
    """
    pos_mask = np.asarray(pos_mask, bool)
    neg_mask = np.asarray(neg_mask, bool)
    B = pos_mask.shape[0]
    sel, thr = [], []
    fn = OHNM_single_image if variant == "model" else OHNM_single_image_pixellink
    for b in range(B):
        n_pos = int(pos_mask[b].sum())
        s, t = fn(neg_conf[b], n_pos, neg_mask[b], ratio)
        sel.append(s)
        thr.append(t)
    return (pos_mask.astype(f32) + np.stack(sel)).astype(f32), np.asarray(thr, f32)


# --------------------------------------------------------------------------- per-pixel terms
ALPHA, GAMMA = 0.25, 2.0


def _term_and_grad(logits2, lab01, term="ce", alpha=ALPHA, gamma=GAMMA):
    """Per-element loss term and d(term)/d(logit of class 1).

    d/d(logit 0) = -d/d(logit 1) for every 2-way softmax term.
    ce    : CE,   grad1 = q1 - lab
    focal : -a_t (1-p_t)^g log p_t  (Lin et al. 2017; a_t = alpha for class 1,
            1-alpha for class 0), grad wrt x_t = a_t (1-p_t)^g (g p_t log p_t - (1-p_t)).
    """
    lab = np.asarray(lab01).astype(np.int64)
    ce = sparse_xent2(logits2, lab)
    q = softmax2(logits2)
    if term == "ce":
        g1 = (q[..., 1] - lab.astype(f32)).astype(f32)
        return ce, g1
    if term == "focal":
        pt = np.take_along_axis(q, lab[..., None], -1)[..., 0].astype(f32)
        logpt = (-ce).astype(f32)
        at = np.where(lab == 1, f32(alpha), f32(1.0 - alpha)).astype(f32)
        om = (f32(1) - pt).astype(f32)
        mod = np.power(om, f32(gamma), dtype=f32)
        val = (-(at * mod * logpt)).astype(f32)
        gt = (at * mod * (f32(gamma) * pt * logpt - om)).astype(f32)  # d/dx_t
        g1 = np.where(lab == 1, gt, -gt).astype(f32)
        return val, g1
    raise ValueError(term)


# --------------------------------------------------------------------------- L5 / L9 / L10
def loss_model(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask=None,
               ratio=3, weight_mode="ohem", term="ce", alpha=ALPHA, gamma=GAMMA):
    """nets/model.py:204-261 ``loss`` (weight_mode='ohem', term='ce'), with
    gradients w.r.t. the two logit tensors.

    weight_mode='pos_only' gives nets/model_vgg_16.py:243-282 ``ohem_loss`` (L9:
    W_pixel = positives only, unguarded pixel normaliser).  term='focal' is the
    L10 ablation (PARITY UNPINNED).  ``training_mask`` is accepted and ignored
    (quirk Q3).  No zero guards on the link normalisers: NaN is data (quirk Q2).
    """
    yp = np.asarray(y_pred_pixel, f32)
    B = yp.shape[0]
    pixel_label = np.asarray(y_true_pixel, f32).reshape(B, -1).astype(np.int32)   # :213 cast truncates
    pixel_pred = yp.reshape(B, -1, 2)                                             # :214
    N = pixel_pred.shape[1]
    pixel_scores = softmax2(pixel_pred)                                           # :216
    pixel_neg_scores = pixel_scores[:, :, 0]                                      # :217
    pos_mask, neg_mask = get_pos_and_neg_masks(pixel_label)                       # :218
    if weight_mode == "ohem":
        M, thr = OHNM_batch(14, pixel_neg_scores, pos_mask, neg_mask, ratio)      # :220
    elif weight_mode == "pos_only":
        M, thr = pos_mask.astype(f32), np.full(B, np.nan, f32)                    # vgg16 :265
    else:
        raise ValueError(weight_mode)
    n_seg_pos = f32(_sum(pos_mask))                                               # :221

    lab_ce = (pixel_label == 1).astype(np.int64)
    pix_term, pix_g1 = _term_and_grad(pixel_pred, lab_ce, term, alpha, gamma)
    s_pix = f32(_sum(pix_term.astype(np.float64) * M))
    if weight_mode == "ohem":
        if n_seg_pos > 0:                                                         # :226-233
            L_pix = _div(s_pix, n_seg_pos)
            pix_scale = _div(2.0, n_seg_pos)
        else:
            L_pix, pix_scale = f32(0), f32(0)
    else:                                                                         # vgg16 :267 unguarded
        L_pix = _div(s_pix, n_seg_pos)
        pix_scale = _div(2.0, n_seg_pos)
    with np.errstate(invalid="ignore"):
        gp1 = (M * pix_scale * pix_g1).astype(f32)
    grad_pixel = np.stack([-gp1, gp1], -1).reshape(yp.shape).astype(f32)

    yl = np.asarray(y_pred_link, f32)
    link_pred = yl.reshape(B, N, 8, 2)
    link_lab = np.asarray(y_true_link, f32).reshape(B, N, 8).astype(np.int32)     # :242
    grad_link = np.zeros((B, N, 8, 2), f32)
    w_link = np.zeros((B, N, 8), f32)
    L_link = np.zeros(8, f32)
    sum_wp = np.zeros(8, f32)
    sum_wn = np.zeros(8, f32)
    s_pos = np.zeros(8, f32)
    s_neg = np.zeros(8, f32)
    for d in range(8):                                                            # :239-254
        lab_d = link_lab[:, :, d]
        lp, ln = get_pos_and_neg_masks(lab_d)
        t, g1 = _term_and_grad(link_pred[:, :, d, :], (lab_d == 1).astype(np.int64), term, alpha, gamma)
        Wp = lp.astype(f32) * M
        Wn = ln.astype(f32) * M
        sum_wp[d] = _sum(Wp)
        sum_wn[d] = _sum(Wn)
        s_pos[d] = _sum(t.astype(np.float64) * Wp)
        s_neg[d] = _sum(t.astype(np.float64) * Wn)
        L_link[d] = _div(s_pos[d], sum_wp[d]) + _div(s_neg[d], sum_wn[d])          # :252-254
        with np.errstate(invalid="ignore", divide="ignore"):
            w = (Wp * _div(1.0, sum_wp[d]) + Wn * _div(1.0, sum_wn[d])).astype(f32)
            gl1 = (w * g1).astype(f32)
        grad_link[:, :, d, 1] = gl1
        grad_link[:, :, d, 0] = -gl1
        w_link[:, :, d] = w
    with np.errstate(invalid="ignore"):
        link_total = f32(np.sum(L_link.astype(np.float64)))                        # :256
        total = f32(link_total + f32(2) * L_pix)                                   # :261
    return dict(loss=total, L_pix=L_pix, L_link=L_link, link_total=link_total,
                n_seg_pos=n_seg_pos, sum_wp=sum_wp, sum_wn=sum_wn,
                s_pix=s_pix, s_pos=s_pos, s_neg=s_neg, thr=thr,
                ohem_mask=M.reshape(yp.shape[:-1]).astype(f32),
                grad_pixel=grad_pixel, grad_link=grad_link.reshape(yl.shape),
                # per-element gradient weights (gradient = weight * (softmax - onehot)): the scale the
                # elementwise tolerance of tests/util.py:grad_close is stated against
                w_pixel=(M * pix_scale).astype(f32).reshape(yp.shape[:-1]),
                w_link=w_link.reshape(yl.shape[:-1] + (8,)))


def ohem_loss_vgg16(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask=None):
    """nets/model_vgg_16.py:243-282 ``ohem_loss`` (L9) — positives-only weights."""
    return loss_model(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask,
                      weight_mode="pos_only")


def cal_link_loss(link_gt, link_pred, W_pixel):
    """nets/model_vgg_16.py:227-241."""
    lab = np.asarray(link_gt, f32).reshape(-1).astype(np.int32)
    pred = np.asarray(link_pred, f32).reshape(-1, 2)
    W = np.asarray(W_pixel, f32).reshape(-1)
    ce = sparse_xent2(pred, (lab == 1).astype(np.int64))
    Wp = (lab == 1).astype(f32) * W
    Wn = (lab == 0).astype(f32) * W
    return f32(_div(_sum(ce.astype(np.float64) * Wp), _sum(Wp)) + _div(_sum(ce.astype(np.float64) * Wn), _sum(Wn)))


# --------------------------------------------------------------------------- L6
def build_loss_pixellink(pixel_cls, link_cls, pixel_labels, link_labels, max_neg_pos_ratio=3):
    """nets/pixellink.py:88-263 ``PixelLinkNet.build_loss``.

    pixel loss = 2 * mean CE (:160,:170); link loss per direction = pos/neg
    balanced CE with zero guards, NOT multiplied by the OHEM mask (:193-212);
    the OHNM mask (:152-156) is diagnostic only (quirk Q6).
    """
    pc = np.asarray(pixel_cls, f32)
    lc = np.asarray(link_cls, f32)
    B, H, W = pc.shape[:3]
    pl = np.asarray(pixel_labels, f32).reshape(B, H, W)
    ll = np.asarray(link_labels, f32).reshape(B, H, W, 8)
    scores = softmax2(pc)
    seg_pos, seg_neg = pl > 0, ~(pl > 0)                                          # :95-104
    sel, thr = OHNM_batch(B, scores[..., 0].reshape(B, -1), seg_pos.reshape(B, -1),
                          seg_neg.reshape(B, -1), max_neg_pos_ratio, variant="pixellink")
    n_seg_pos = f32(_sum(sel))                                                    # :155 (sum of the SELECTED mask)
    lab = seg_pos.astype(np.int64)
    ce = sparse_xent2(pc, lab)
    cnt = f32(B * H * W)
    pixel_cls_loss = f32(_sum(ce) / float(cnt))                                   # :160
    q = softmax2(pc)
    gp1 = ((q[..., 1] - lab.astype(f32)) * _div(2.0, cnt)).astype(f32)
    grad_pixel = np.stack([-gp1, gp1], -1).astype(f32)

    lpos = ll > 0
    link_pred = lc.reshape(B, H, W, 8, 2)
    grad_link = np.zeros_like(link_pred)
    L = np.zeros(8, f32)
    pos_n = np.zeros(8, f32)
    neg_n = np.zeros(8, f32)
    for d in range(8):                                                            # :182-212
        labd = lpos[..., d].astype(np.int64)
        ced = sparse_xent2(link_pred[..., d, :], labd)
        qd = softmax2(link_pred[..., d, :])
        pw = lpos[..., d].astype(f32)
        nw = (~lpos[..., d]).astype(f32)
        pos_n[d], neg_n[d] = _sum(pw), _sum(nw)
        ps = _div(1.0, pos_n[d]) if pos_n[d] != 0 else f32(0)                     # :198-211 guards
        ns = _div(1.0, neg_n[d]) if neg_n[d] != 0 else f32(0)
        L[d] = f32(_sum(ced.astype(np.float64) * pw) * float(ps) + _sum(ced.astype(np.float64) * nw) * float(ns))
        w = (pw * ps + nw * ns).astype(f32)
        g1 = (w * (qd[..., 1] - labd.astype(f32))).astype(f32)
        grad_link[..., d, 1] = g1
        grad_link[..., d, 0] = -g1
    link_total = f32(np.sum(L.astype(np.float64)))                                 # :253
    return dict(pixel_cls_loss=pixel_cls_loss, losses=[f32(2) * pixel_cls_loss, link_total],
                loss=f32(f32(2) * pixel_cls_loss + link_total), L_link=L,
                pos_n=pos_n, neg_n=neg_n, n_seg_pos=n_seg_pos, thr=thr,
                ohem_mask=sel.reshape(B, H, W).astype(f32),
                grad_pixel=grad_pixel, grad_link=grad_link.reshape(lc.shape))


# --------------------------------------------------------------------------- L7 / L8
def dice_coefficient(y_true_cls, y_pred_cls, training_mask, with_grad=False):
    """nets/model.py:145-159 == nets/model_vgg_16.py:179-193.

    One scalar over the WHOLE tensor.  grad wrt pred = -2 m (t U - I) / U^2.
    """
    t = np.asarray(y_true_cls, f32)
    p = np.asarray(y_pred_cls, f32)
    m = np.asarray(training_mask, f32)
    eps = 1e-5
    I = f32(_sum((t * p * m).astype(np.float64)))
    U = f32(f32(_sum((t * m).astype(np.float64))) + f32(_sum((p * m).astype(np.float64))) + f32(eps))
    loss = f32(f32(1.0) - f32(2) * I / U)
    if not with_grad:
        return loss
    g = (f32(-2) * m * (t * U - I) / (U * U)).astype(f32)
    return loss, g, I, U


def loss_vgg16_dice(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask):
    """nets/model_vgg_16.py:196-225 — 2*dice(pixel) + sum_d dice(link_d); the
    predictions are PROBABILITIES (sigmoid outputs, model_vgg_16.py:129-131)."""
    tp = np.asarray(y_true_pixel, f32)
    pp = np.asarray(y_pred_pixel, f32)
    tl = np.asarray(y_true_link, f32)
    plk = np.asarray(y_pred_link, f32)
    m = np.asarray(training_mask, f32)
    lp, gp, I0, U0 = dice_coefficient(tp, pp, m, True)
    total = f32(2) * lp
    grad_link = np.zeros_like(plk)
    Ls = np.zeros(9, f32)
    Is = np.zeros(9, f32)
    Us = np.zeros(9, f32)
    Ls[0], Is[0], Us[0] = lp, I0, U0
    link_loss = f32(0)
    for d in range(8):
        ld, gd, I, U = dice_coefficient(tl[..., d:d + 1], plk[..., d:d + 1], m, True)
        link_loss = f32(link_loss + ld)
        grad_link[..., d:d + 1] = gd
        Ls[d + 1], Is[d + 1], Us[d + 1] = ld, I, U
    return dict(loss=f32(link_loss + total), dice=Ls, I=Is, U=Us,
                grad_pixel=(f32(2) * gp).astype(f32), grad_link=grad_link)
