"""Mirror of the head functions of the reference's ``nets/model.py`` (lines 145-261).

Same names, same positional signatures.  Inputs may be numpy arrays (host round
trip, like a ``tf.py_func``) or CUDA ``torch.Tensor`` s (zero copy, autograd-aware:
``loss(...).backward()`` fills ``.grad`` of the two logit tensors with the gradients
the fused kernel produced in the same pass).  Everything executes in libplhead.so.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, head

__all__ = ["feature_fusion", "dice_coefficient", "OHNM_single_image", "OHNM_batch", "get_pos_and_neg_masks", "loss",
           "loss_with_stats"]


class _FusedLoss(torch.autograd.Function):
    """loss value + analytic gradients from one fused launch sequence."""

    @staticmethod
    def forward(ctx, y_pred_pixel, y_pred_link, y_true_pixel, y_true_link, cfg):
        need = y_pred_pixel.requires_grad or y_pred_link.requires_grad
        out = head.pixellink_loss_raw(y_pred_pixel.detach(), y_pred_link.detach(), y_true_pixel, y_true_link, cfg,
                                      want_grad=need)
        ctx.need = need
        if need:
            ctx.save_for_backward(out["grad_pixel"], out["grad_link"])
        ctx.stats = out["stats"]
        return out["stats"][_lib.ST_TOTAL].clone()

    @staticmethod
    def backward(ctx, g):
        gp, gl = ctx.saved_tensors
        return g * gp, g * gl, None, None, None


def _prep(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link):
    yp, np_in = head.to_device(y_pred_pixel)
    yl, _ = head.to_device(y_pred_link, device=yp.device)
    tp, _ = head.to_device(y_true_pixel, device=yp.device)
    tl, _ = head.to_device(y_true_link, device=yp.device)
    return tp, yp, tl, yl, np_in


def loss_with_stats(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask=None,
                    cfg: head.LossConfig = head.LossConfig(), want_grad=True, want_mask=True):
    """Like :func:`loss` but returns the whole result dict (stats, gradients, OHEM mask)."""
    tp, yp, tl, yl, np_in = _prep(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link)
    out = head.pixellink_loss_raw(yp, yl, tp, tl, cfg, want_grad=want_grad, want_mask=want_mask)
    if np_in:
        return {k: v.cpu().numpy() for k, v in out.items()}
    return out


def loss(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask):
    """nets/model.py:204-261.  Returns ``weight_link_loss + 2 * classification_loss``.

    ``training_mask`` is accepted and ignored, exactly like the reference (quirk Q3).
    The reference's literal ``OHNM_batch(14, ...)`` (model.py:220) is generalised to
    the actual batch size (quirk Q1).  NaN from an empty link class is returned as
    data (quirk Q2).
    """
    tp, yp, tl, yl, np_in = _prep(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link)
    if np_in:
        out = head.pixellink_loss_raw(yp, yl, tp, tl, head.LossConfig(), want_grad=False)
        return out["stats"][_lib.ST_TOTAL].cpu().numpy()[()]
    return _FusedLoss.apply(yp, yl, tp, tl, head.LossConfig())


def get_pos_and_neg_masks(labels):
    """nets/model.py:199-202."""
    if isinstance(labels, torch.Tensor):
        return labels == 1, labels == 0
    labels = np.asarray(labels)
    return labels == 1, labels == 0


def _check_scores(sc):
    """The selection kernel ranks the fp32 bit patterns of the scores, which orders them only on [0, 1] (they are
    softmax probabilities in the reference, nets/model.py:216-217).  Anything else — negative, above 1, NaN —
    would silently give a wrong threshold, so it is rejected here (one reduction; the fused loss computes its own
    scores and does not come through this check)."""
    lo, hi = torch.aminmax(sc)
    if not (float(lo) >= 0.0 and float(hi) <= 1.0):     # NaN fails both comparisons
        raise ValueError("OHNM scores must be probabilities in [0, 1] (got min %r, max %r)" % (float(lo), float(hi)))


def OHNM_batch(batch_size, neg_conf, pos_mask, neg_mask):
    """nets/model.py:186-197 — ``float(pos_mask) + selected_neg_mask`` with the per-image
    radix-select kernel.  ``batch_size`` is accepted; the real batch is ``pos_mask.shape[0]``."""
    sc, np_in = head.to_device(neg_conf)
    _check_scores(sc)
    shape = tuple(sc.shape)
    B = shape[0]
    pm, _ = head.to_device(pos_mask, torch.uint8, sc.device)
    nm, _ = head.to_device(neg_mask, torch.uint8, sc.device)
    sel, _thr = head.ohnm_batch_raw(sc.reshape(B, -1), pm.reshape(B, -1), nm.reshape(B, -1))
    sel = sel.reshape(shape)
    return sel.cpu().numpy() if np_in else sel


def OHNM_single_image(scores, n_pos, neg_mask):
    """nets/model.py:161-184 — mask of the selected negatives of ONE image."""
    sc, np_in = head.to_device(scores)
    _check_scores(sc)
    shape = tuple(sc.shape)
    nm, _ = head.to_device(neg_mask, torch.uint8, sc.device)
    npos = torch.as_tensor([int(n_pos)], dtype=torch.int32, device=sc.device)
    pm = torch.zeros_like(nm)
    sel, _thr = head.ohnm_batch_raw(sc.reshape(1, -1), pm.reshape(1, -1), nm.reshape(1, -1), n_pos=npos)
    sel = sel.reshape(shape)
    return sel.cpu().numpy() if np_in else sel


class _Dice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, mask):
        outv, grad = head.dice_raw(y_true, y_pred.detach(), mask, want_grad=y_pred.requires_grad)
        if y_pred.requires_grad:
            ctx.save_for_backward(grad)
        return outv[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g * grad, None, None


def dice_coefficient(y_true_cls, y_pred_cls, training_mask):
    """nets/model.py:145-159 — one dice scalar over the whole tensor."""
    p, np_in = head.to_device(y_pred_cls)
    t, _ = head.to_device(y_true_cls, device=p.device)
    m, _ = head.to_device(training_mask, device=p.device)
    if m.numel() != p.numel():
        m = m.expand_as(p).contiguous()
    if np_in:
        outv, _ = head.dice_raw(t, p, m, want_grad=False)
        return outv[0].cpu().numpy()[()]
    return _Dice.apply(p, t, m)


def feature_fusion(feature_maps, params):
    """nets/model.py:129-141 (with ``unpool`` :14-15): the fusion that ends ``model()``,

        pixel_1 = unpool(c(f0)) + c(f1);  pixel_2 = unpool(pixel_1) + c(f2);  pixel_3 = unpool(pixel_2) + c(f3)
        pixel_4 = conv1x1(pixel_3)                        (and the same for the 16 link channels)

    where ``c`` is ``slim.conv2d(., n, 1)`` under the arg_scope of :103-107 — batch norm, here in its INFERENCE
    form folded to a per-channel scale and shift, then ReLU.  ``feature_maps`` = [pool5, pool4, pool3, pool2]
    (NHWC, each level twice the previous one); ``params["pixel"]`` / ``params["link"]`` = five
    ``(weights [K,n], scale [n] | None, shift [n])`` triples: the four fuse convolutions f0..f3 and the plain last one.
    Returns ``(pixel_4 [B,H,W,2], link_4 [B,H,W,16])`` (numpy in -> numpy out)."""
    import torch
    np_in, dev, x = False, None, []
    for f in feature_maps:
        t, was_np = head.to_device(f, device=dev)
        dev, np_in = t.device, np_in or was_np
        x.append(t)

    def cat(i, j, ones=False):
        a, b = params["pixel"][i][j], params["link"][i][j]
        if a is None and b is None:
            return None
        a = np.ones(2, np.float32) if a is None else a
        b = np.ones(16, np.float32) if b is None else b
        ta, _ = head.to_device(a, device=dev)
        tb, _ = head.to_device(b, device=dev)
        return torch.cat([ta, tb], dim=-1).contiguous()

    prev = None
    for i in range(4):
        feat = (x[i], cat(i, 0), cat(i, 1), cat(i, 2), True)
        if i < 3:
            prev = head.head_fuse_level_raw([feat], prev=prev)
        else:
            w_out = torch.zeros((18, 18), dtype=torch.float32, device=dev)
            w_out[:2, :2], _ = head.to_device(params["pixel"][4][0], device=dev)
            w_out[2:, 2:], _ = head.to_device(params["link"][4][0], device=dev)
            if params["pixel"][4][1] is not None or params["link"][4][1] is not None:
                w_out = w_out * cat(4, 1)[None, :]
            pixel_4, link_4 = head.head_fuse_level_raw([feat], prev=prev, w_out=w_out, b_out=cat(4, 2))
    if np_in:
        return pixel_4.cpu().numpy(), link_4.cpu().numpy()
    return pixel_4, link_4
