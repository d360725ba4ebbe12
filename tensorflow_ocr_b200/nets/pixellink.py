"""Mirror of the head part of the reference's ``nets/pixellink.py`` (lines 69-72, 88-263)."""
from __future__ import annotations

import torch

from .. import _lib, head
from .model import _FusedLoss

__all__ = ["PixelLinkNet"]


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
