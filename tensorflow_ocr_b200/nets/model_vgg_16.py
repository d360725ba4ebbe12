"""Mirror of the head functions of the reference's ``nets/model_vgg_16.py`` (lines 179-282)."""
from __future__ import annotations

import torch

from .. import _lib, head
from .model import _FusedLoss, _prep, dice_coefficient  # noqa: F401  (dice_coefficient :179-193 is identical)

__all__ = ["dice_coefficient", "loss", "cal_link_loss", "ohem_loss"]

_POS_ONLY = head.LossConfig(variant=_lib.VARIANT_POS_ONLY)


class _DiceHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_pix, p_link, t_pix, t_link, mask):
        need = p_pix.requires_grad or p_link.requires_grad
        outv, gp, gl = head.dice_head_raw(t_pix, p_pix.detach(), t_link, p_link.detach(), mask, want_grad=need)
        if need:
            ctx.save_for_backward(gp, gl)
        return outv[0].clone()

    @staticmethod
    def backward(ctx, g):
        gp, gl = ctx.saved_tensors
        return g * gp, g * gl, None, None, None


def loss(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask):
    """nets/model_vgg_16.py:196-225 — 2*dice(pixel) + sum of the 8 link dice terms.
    Predictions are probabilities ([B,H,W,1] and [B,H,W,8])."""
    pp, np_in = head.to_device(y_pred_pixel)
    pl, _ = head.to_device(y_pred_link, device=pp.device)
    tp, _ = head.to_device(y_true_pixel, device=pp.device)
    tl, _ = head.to_device(y_true_link, device=pp.device)
    m, _ = head.to_device(training_mask, device=pp.device)
    if np_in:
        outv, _, _ = head.dice_head_raw(tp, pp, tl, pl, m, want_grad=False)
        return outv[0].cpu().numpy()[()]
    return _DiceHead.apply(pp, pl, tp, tl, m)


def ohem_loss(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link, training_mask):
    """nets/model_vgg_16.py:243-282 — despite the name: positives-only pixel weight, no mining."""
    tp, yp, tl, yl, np_in = _prep(y_true_pixel, y_pred_pixel, y_true_link, y_pred_link)
    if np_in:
        out = head.pixellink_loss_raw(yp, yl, tp, tl, _POS_ONLY, want_grad=False)
        return out["stats"][_lib.ST_TOTAL].cpu().numpy()[()]
    return _FusedLoss.apply(yp, yl, tp, tl, _POS_ONLY)


def cal_link_loss(link_gt, link_pred, W_pixel):
    """nets/model_vgg_16.py:227-241 — one direction's pos/neg balanced CE with pixel weight W.

    Runs the fused kernel with this direction in slot 0 and the weight as the pixel
    label map (W_pixel is 0/1 in the reference: `tf.equal(y_true_pixel, 1)` cast to float).
    """
    lp, np_in = head.to_device(link_pred)
    dev = lp.device
    lg, _ = head.to_device(link_gt, device=dev)
    w, _ = head.to_device(W_pixel, device=dev)
    M = lg.numel()
    link_logits = torch.zeros((1, 1, M, 16), dtype=torch.float32, device=dev)
    link_logits[..., 0:2] = lp.reshape(1, 1, M, 2)
    link_lab = torch.zeros((1, 1, M, 8), dtype=torch.float32, device=dev)
    link_lab[..., 0] = lg.reshape(1, 1, M)
    pix_logits = torch.zeros((1, 1, M, 2), dtype=torch.float32, device=dev)
    out = head.pixellink_loss_raw(pix_logits, link_logits, w.reshape(1, 1, M, 1).contiguous(), link_lab, _POS_ONLY,
                                  want_grad=False)
    val = out["stats"][_lib.ST_L_LINK]
    return val.cpu().numpy()[()] if np_in else val.clone()
