"""Mirror of the head part of the reference's ``nets/pixellink.py``: the feature fusion that produces the logits
(lines 37-38, 56-67) and the loss (lines 69-72, 88-263)."""
from __future__ import annotations

import torch

from .. import _lib, head
from .model import _FusedLoss

__all__ = ["PixelLinkNet", "pixellink_layers"]


class PixelLinkNet(object):
    """Carries ``pixel_cls`` / ``link_cls`` logits like the reference object
    (nets/pixellink.py:69-72) and exposes ``build_loss`` with the same signature.

    The backbone (nets/pixellink.py:8-67) is out of scope; construct with the two
    logit tensors the backbone would have produced.  ``max_neg_pos_ratio`` and
    ``batch_size_per_gpu`` live in the reference's missing ``config`` module; they are
    constructor arguments here (upstream pixel_link default 3).
    """

    def __init__(self, pixel_cls, link_cls, max_neg_pos_ratio=3):
        self.pixel_cls, self._np_in = head.to_device(pixel_cls)
        self.link_cls, _ = head.to_device(link_cls, device=self.pixel_cls.device)
        self.max_neg_pos_ratio = max_neg_pos_ratio
        self.losses = []          # stands in for tf.GraphKeys.LOSSES (:170, :254)
        self.summaries = {}

    def build_loss(self, pixel_labels, link_labels, do_summary=True):
        """nets/pixellink.py:88-263.  Appends ``[2*pixel_cls_loss, link_total_loss]`` to
        ``self.losses`` (the reference adds exactly these two to tf.GraphKeys.LOSSES) and
        returns None like the reference.  The OHNM mask is diagnostic only (quirk Q6) and
        is exposed through ``self.summaries``."""
        cfg = head.LossConfig(variant=_lib.VARIANT_PIXELLINK, neg_pos_ratio=self.max_neg_pos_ratio)
        dev = self.pixel_cls.device
        pl, _ = head.to_device(pixel_labels, device=dev)
        ll, _ = head.to_device(link_labels, device=dev)
        if self.pixel_cls.requires_grad or self.link_cls.requires_grad:
            total = _FusedLoss.apply(self.pixel_cls, self.link_cls, pl, ll, cfg)
            stats = total.grad_fn.stats if total.grad_fn is not None else None
            self.total_loss = total
        else:
            out = head.pixellink_loss_raw(self.pixel_cls, self.link_cls, pl, ll, cfg, want_grad=False, want_mask=True)
            stats = out["stats"]
            self.total_loss = stats[_lib.ST_TOTAL]
            self.summaries["seg_selected_mask"] = out["ohem_mask"]
        if stats is not None:
            self.losses = [2.0 * stats[_lib.ST_L_PIX], stats[_lib.ST_LINK_TOTAL]]
            if do_summary:
                self.summaries["pixel_cls_loss"] = stats[_lib.ST_L_PIX]
                self.summaries["n_seg_pos"] = stats[_lib.ST_N_SEG_POS]
                self.summaries["link_weighted_loss"] = stats[_lib.ST_L_LINK:_lib.ST_L_LINK + 8]
        if self._np_in:
            self.losses = [l.cpu().numpy()[()] for l in self.losses]
        return None


_FUSE_SCOPES = (("fc7", 6), ("conv5_3", 5), ("conv4_3", 4), ("conv3_3", 3))


def _cat18(pix, link, device):
    """Pixel (2 columns) and link (16 columns) parameters side by side: the two branches share one pass."""
    a, _ = head.to_device(pix, device=device)
    b, _ = head.to_device(link, device=device)
    return torch.cat([a, b], dim=-1).contiguous()


def pixellink_layers(end_points, params, decode_config=None):
    """nets/pixellink.py:56-67 ``_add_pixellink_layers`` (with ``unpool`` :37-38) on the GPU:

        s1 = conv1x1(fc7) + conv1x1(conv5_3);  s2 = unpool(s1) + conv1x1(conv4_3);
        s3 = unpool(s2) + conv1x1(conv3_3);    pixel_cls / link_cls = conv1x1(s3)

    ``end_points``: the NHWC feature maps ``fc7``, ``conv5_3`` (1/16), ``conv4_3`` (1/8), ``conv3_3`` (1/4);
    ``params[scope] = (weights [K,n], biases [n])`` for the reference's ten variable scopes
    ``stage_{6,5,4,3}_{pixel,link}_fuse``, ``text_predication``, ``link_predication``.
    Returns ``(pixel_cls [B,H,W,2], link_cls [B,H,W,16])`` (numpy in -> numpy out).  With ``decode_config`` (a
    ``head.DecodeConfig``) a third element: the decode's threshold words of these logits (what
    ``head.decode_flags_raw`` would compute), for ``head.decode_from_flags_raw`` — the decode then never reads the
    logits."""
    x = {}
    np_in = False
    dev = None
    for name, _ in _FUSE_SCOPES:
        x[name], was_np = head.to_device(end_points[name], device=dev)
        dev = x[name].device
        np_in = np_in or was_np
    w_out = torch.zeros((18, 18), dtype=torch.float32, device=dev)     # the two last convolutions as one block matrix
    w_out[:2, :2], _ = head.to_device(params["text_predication"][0], device=dev)
    w_out[2:, 2:], _ = head.to_device(params["link_predication"][0], device=dev)
    b_out = _cat18(params["text_predication"][1], params["link_predication"][1], dev)
    feats = {}
    for name, st in _FUSE_SCOPES:
        w = _cat18(params["stage_%d_pixel_fuse" % st][0], params["stage_%d_link_fuse" % st][0], dev)
        b = _cat18(params["stage_%d_pixel_fuse" % st][1], params["stage_%d_link_fuse" % st][1], dev)
        feats[name] = (x[name], w, None, b, False)
    # The fuse convolutions carry no activation, so the last 1x1 convolutions commute with everything after stage 2:
    #   (unpool(s2) + x3 W3 + b3) W_out + b_out = unpool(s2 W_out) + x3 (W3 W_out) + (b3 W_out + b_out).
    # W_out is applied to s2 (a quarter of the pixels) and folded into the conv3_3 weights; the largest level then
    # has no per-pixel output matrix left.
    w3 = (feats["conv3_3"][1].double() @ w_out.double()).float().contiguous()
    b3 = (feats["conv3_3"][3].double() @ w_out.double() + b_out.double()).float().contiguous()
    s1 = head.head_fuse_level_raw([feats["fc7"], feats["conv5_3"]])
    s2 = head.head_fuse_level_raw([feats["conv4_3"]], prev=s1, w_out=w_out, logits=False)
    res = head.head_fuse_level_raw([(x["conv3_3"], w3, None, b3, False)], prev=s2, logits=True, flags_cfg=decode_config)
    return tuple(t.cpu().numpy() for t in res) if np_in else res
