"""Host-side engine: thin, allocation-caching wrappers that hand device pointers to
the C ABI (include/plhead.h).  torch is used for device memory, streams and
host<->device copies only; every arithmetic step of the head runs in libplhead.so.

All functions here are asynchronous with respect to the host (they enqueue on
torch's current CUDA stream and return device tensors) unless the caller passed
numpy arrays, in which case the results are copied back (tf.py_func-style,
tool/pixellink_fn.py:114,157).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

__all__ = ["DecodeConfig", "LossConfig", "pixellink_loss_raw", "decode_raw", "loss_and_decode_raw",
           "ohnm_batch_raw", "dice_raw", "dice_head_raw", "east_loss_raw", "restore_rectangle_raw",
           "pixel_detect_raw", "min_area_boxes_raw", "lanms_raw", "contour_boxes_raw", "to_device", "launch_count"]


# ----------------------------------------------------------------------------- configuration
@dataclass(frozen=True)
class LossConfig:
    """Constants the reference hard-codes in the head (SURVEY.md §5 "Config / flags")."""
    variant: int = _lib.VARIANT_MODEL
    term: int = _lib.TERM_CE
    neg_pos_ratio: int = 3        # nets/model.py:171
    focal_alpha: float = 0.25     # Lin et al. 2017 (not in the reference)
    focal_gamma: float = 2.0
    main_only: bool = False       # measurement hook: rerun the main pass on an already prepared workspace
    split_counts: bool = False    # scheduling hint: normalisers in their own pass instead of fused into the selection
    chain_pdl: bool = False       # scheduling hint: the stream's preceding kernel is this library's (early launch ok)

    def c_struct(self):
        p = _lib.LossParams(self.variant, self.term, self.neg_pos_ratio, self.focal_alpha, self.focal_gamma)
        p.reserved[0] = (1 if self.main_only else 0) | (2 if self.split_counts else 0) | (4 if self.chain_pdl else 0)
        return p


@dataclass(frozen=True)
class DecodeConfig:
    pixel_thresh: float = 0.8               # test_pixellink_fast.py:12
    link_thresh: float = 0.9                # test_pixellink_fast.py:13
    min_size: int = 10                      # test_pixellink_fast.py:174 (200 in test_pixellink.py:177)
    scale: Tuple[float, float] = (4.0, 3.75)  # (1280/320, 720/192) test_pixellink_fast.py:196-197
    max_boxes: int = 128
    phase: int = 0    # plh_decode_params.reserved[0]: 0 whole decode, 4 tile pass only, 8 resume after the tile pass
    form: str = "auto"  # "auto" = "tiled" (5 launches, forest in L2) | "resident" (one 8-CTA cluster per image, map in shared memory, 3 launches): same results

    def c_struct(self):
        p = _lib.DecodeParams(self.pixel_thresh, self.link_thresh, self.min_size, self.max_boxes,
                              float(self.scale[0]), float(self.scale[1]))
        p.reserved[0] = self.phase | {"auto": 0, "tiled": 16, "resident": 32}[self.form]
        return p


# ----------------------------------------------------------------------------- plumbing
def _require_gpu(device: Optional[torch.device] = None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("tensorflow_ocr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    major, _ = torch.cuda.get_device_capability(dev)
    if major != 10:
        raise RuntimeError("libplhead.so is built for sm_100a only; device %s has capability %d.x" % (dev, major))
    return dev


def to_device(x, dtype=torch.float32, device=None):
    """numpy / torch (any device) -> contiguous CUDA tensor of `dtype`.  Returns (tensor, was_numpy)."""
    was_numpy = not isinstance(x, torch.Tensor)
    if was_numpy:
        dev = _require_gpu(device)
        t = torch.as_tensor(np.ascontiguousarray(x))
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.to(dev, non_blocking=True), True
    t = x
    if not t.is_cuda:
        t = t.to(_require_gpu(device))
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous(), False


_ws_cache = {}      # key -> [tensor, pinned]; insertion order = age
_WS_CACHE_MAX = 64


def _workspace(op: int, B: int, H: int, W: int, K: int, device: torch.device) -> torch.Tensor:
    """Caller-owned workspace, cached per (device, stream, op, shape): the library keeps no state.

    A workspace handed out while the stream is being captured into a CUDA graph is baked into that graph
    and is therefore pinned for the life of the process; only unpinned entries are ever evicted (oldest
    first), so a replayed graph never points at freed memory."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, op, B, H, W, K)
    capturing = torch.cuda.is_current_stream_capturing()
    ent = _ws_cache.get(key)
    if ent is None:
        n = _lib.load().plh_workspace_bytes(op, B, H, W, K)
        if n == 0:
            raise ValueError("bad shape for workspace query: op=%d B=%d H=%d W=%d" % (op, B, H, W))
        if len(_ws_cache) >= _WS_CACHE_MAX:
            for k in [k for k, e in _ws_cache.items() if not e[1]][: len(_ws_cache) - _WS_CACHE_MAX + 1]:
                del _ws_cache[k]
        ent = _ws_cache[key] = [torch.empty(n, dtype=torch.uint8, device=device), capturing]
    elif capturing:
        ent[1] = True
    return ent[0]


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count() -> int:
    return int(_lib.load().plh_launch_count())


# ----------------------------------------------------------------------------- loss
def _check_head_shapes(pix_logits, link_logits, pix_lab, link_lab):
    if pix_logits.dim() != 4 or pix_logits.shape[-1] != 2:
        raise ValueError("pixel logits must be [B,H,W,2], got %s" % (tuple(pix_logits.shape),))
    B, H, W, _ = pix_logits.shape
    if tuple(link_logits.shape) != (B, H, W, 16):
        raise ValueError("link logits must be [B,H,W,16], got %s" % (tuple(link_logits.shape),))
    if pix_lab is not None and pix_lab.numel() != B * H * W:
        raise ValueError("pixel labels must have B*H*W elements, got %s" % (tuple(pix_lab.shape),))
    if link_lab is not None and tuple(link_lab.shape) != (B, H, W, 8):
        raise ValueError("link labels must be [B,H,W,8], got %s" % (tuple(link_lab.shape),))
    return B, H, W


def pixellink_loss_raw(pix_logits, link_logits, pix_lab, link_lab, cfg: LossConfig = LossConfig(),
                       want_grad: bool = True, want_mask: bool = False,
                       decode: Optional[DecodeConfig] = None, out: Optional[dict] = None,
                       workspace: Optional[torch.Tensor] = None) -> dict:
    """Fused PixelLink loss fwd+bwd on CUDA tensors (plh_pixellink_loss).

    Returns device tensors: stats [64+B], grad_pixel, grad_link (if want_grad),
    ohem_mask uint8 [B,H,W] (if want_mask), flags uint16 [B,H,W] (if decode is given).
    `out` may carry preallocated tensors of those names to be reused.
    """
    lib = _lib.load()
    B, H, W = _check_head_shapes(pix_logits, link_logits, pix_lab, link_lab)
    dev = pix_logits.device
    _require_gpu(dev)
    out = {} if out is None else out

    def buf(name, shape, dtype):
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t

    stats = buf("stats", (_lib.STATS_FLOATS + B,), torch.float32)
    gp = buf("grad_pixel", (B, H, W, 2), torch.float32) if want_grad else None
    gl = buf("grad_link", (B, H, W, 16), torch.float32) if want_grad else None
    mask = buf("ohem_mask", (B, H, W), torch.uint8) if want_mask else None
    flags = buf("flags", (B, H, W), torch.int16) if decode is not None else None
    ws = workspace if workspace is not None else _workspace(_lib.OP_LOSS, B, H, W, 0, dev)
    lp = cfg.c_struct()
    dp = decode.c_struct() if decode is not None else None
    with torch.cuda.device(dev):
        rc = lib.plh_pixellink_loss(_p(pix_logits), _p(link_logits), _p(pix_lab), _p(link_lab), None, B, H, W,
                                    C.byref(lp), _p(stats), _p(gp), _p(gl), _p(mask), _p(flags),
                                    C.byref(dp) if dp is not None else None, _p(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "plh_pixellink_loss")
    return out


def ohnm_batch_raw(scores, pos_mask, neg_mask, variant=_lib.VARIANT_MODEL, ratio=3, n_pos=None):
    """plh_ohnm_batch: scores [B,N] fp32, masks [B,N] uint8 -> (selected fp32 [B,N], thr [B])."""
    lib = _lib.load()
    B, N = scores.shape
    dev = scores.device
    _require_gpu(dev)
    sel = torch.empty((B, N), dtype=torch.float32, device=dev)
    thr = torch.empty((B,), dtype=torch.float32, device=dev)
    ws = torch.empty(lib.plh_workspace_bytes(_lib.OP_LOSS, B, 1, N, 0), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.plh_ohnm_batch(_p(scores), _p(pos_mask), _p(neg_mask), _p(n_pos), B, N, variant, ratio, _p(sel),
                                _p(thr), _p(ws), 0 if ws is None else ws.numel(), _stream(dev))
    _lib.check(rc, "plh_ohnm_batch")
    return sel, thr


# ----------------------------------------------------------------------------- decode
def _decode_outputs(B, H, W, K, dev, out, want_rects):
    out = {} if out is None else out

    def buf(name, shape, dtype):
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=dev)
            out[name] = t
        return t

    labels = buf("labels", (B, H, W), torch.int32)
    boxes = buf("boxes", (B, K, 4, 2), torch.int32)
    n_boxes = buf("n_boxes", (B,), torch.int32)
    comp = buf("comp", (B, K, 2), torch.int32)
    rects = buf("rects", (B, K, 5), torch.float32) if want_rects else None
    return out, labels, boxes, n_boxes, comp, rects


def decode_raw(pix_logits, link_logits, cfg: DecodeConfig = DecodeConfig(), out: Optional[dict] = None,
               want_rects: bool = True, workspace: Optional[torch.Tensor] = None) -> dict:
    """plh_decode on CUDA tensors.  Returns labels int32 [B,H,W], boxes int32 [B,K,4,2],
    n_boxes int32 [B], comp int32 [B,K,2] (label, size), rects fp32 [B,K,5]."""
    lib = _lib.load()
    B, H, W = _check_head_shapes(pix_logits, link_logits, None, None)
    dev = pix_logits.device
    _require_gpu(dev)
    K = cfg.max_boxes
    out, labels, boxes, n_boxes, comp, rects = _decode_outputs(B, H, W, K, dev, out, want_rects)
    ws = workspace if workspace is not None else _workspace(_lib.OP_DECODE, B, H, W, K, dev)
    dp = cfg.c_struct()
    with torch.cuda.device(dev):
        rc = lib.plh_decode(_p(pix_logits), _p(link_logits), B, H, W, C.byref(dp), _p(labels), _p(boxes),
                            _p(n_boxes), _p(rects), _p(comp), _p(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "plh_decode")
    return out


def decode_from_flags_raw(flags, cfg: DecodeConfig = DecodeConfig(), out: Optional[dict] = None,
                          want_rects: bool = True) -> dict:
    lib = _lib.load()
    B, H, W = flags.shape
    dev = flags.device
    K = cfg.max_boxes
    out, labels, boxes, n_boxes, comp, rects = _decode_outputs(B, H, W, K, dev, out, want_rects)
    ws = _workspace(_lib.OP_DECODE, B, H, W, K, dev)
    dp = cfg.c_struct()
    with torch.cuda.device(dev):
        rc = lib.plh_decode_from_flags(_p(flags), B, H, W, C.byref(dp), _p(labels), _p(boxes), _p(n_boxes),
                                       _p(rects), _p(comp), _p(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "plh_decode_from_flags")
    return out


def decode_flags_raw(pix_logits, link_logits, cfg: DecodeConfig = DecodeConfig(), out: Optional[dict] = None) -> dict:
    """plh_decode_flags: the threshold pass alone -> out["flags"] uint16 [B,H,W]."""
    lib = _lib.load()
    B, H, W = pix_logits.shape[:3]
    dev = pix_logits.device
    _require_gpu(dev)
    out = {} if out is None else out
    flags = out.get("flags")
    if flags is None or flags.shape != (B, H, W):
        flags = out["flags"] = torch.empty((B, H, W), dtype=torch.int16, device=dev)
    dp = cfg.c_struct()
    with torch.cuda.device(dev):
        rc = lib.plh_decode_flags(_p(pix_logits), _p(link_logits), B, H, W, C.byref(dp), _p(flags), _stream(dev))
    _lib.check(rc, "plh_decode_flags")
    return out


_aux_streams = {}


def _aux_stream(dev: torch.device) -> torch.cuda.Stream:
    """The second stream of the head step, one per (device, calling stream): two callers that drive the step on
    two streams of one device (batches in flight) must not share their decode branch."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    st = _aux_streams.get(key)
    if st is None:
        st = torch.cuda.Stream(dev)
        _aux_streams[key] = st
    return st


def loss_and_decode_raw(pix_logits, link_logits, pix_lab, link_lab, lcfg: LossConfig = LossConfig(),
                        dcfg: DecodeConfig = DecodeConfig(), out: Optional[dict] = None,
                        want_rects: bool = False, parallel: bool = True, schedule: Optional[str] = None) -> dict:
    """The head step: loss fwd+bwd and decode of the same logits.

    parallel=True (default): the two pipelines are independent chains of kernels and run on two streams
    (fork/join with events; under CUDA-graph capture this becomes two parallel branches).
      schedule "fork" (default): both chains start at once — decode (flags, components, boxes) on
        the auxiliary stream, loss (selection, main pass) on the caller's.
      schedule "tile_first": the round-1 interleaving for the tiled decode — its threshold + tile-labelling
        kernel runs first and alone, the selection kernel launches programmatically under its tail.
    parallel=False: one stream; the loss kernel emits the 2 B/px threshold flags and the decode
    starts from them (one read of the logits instead of two).
    """
    dev = pix_logits.device
    if not parallel:
        out = pixellink_loss_raw(pix_logits, link_logits, pix_lab, link_lab, lcfg, True, False, dcfg, out)
        return decode_from_flags_raw(out["flags"], dcfg, out, want_rects)
    out = {} if out is None else out
    cur = torch.cuda.current_stream(dev)
    aux = _aux_stream(dev)
    B, H, W = pix_logits.shape[:3]
    with torch.cuda.device(dev):
        ws = _workspace(_lib.OP_DECODE, B, H, W, dcfg.max_boxes, dev)
    if schedule is None:
        schedule = "fork" if dcfg.form == "resident" else "tile_first"
    if schedule == "fork":
        aux.wait_stream(cur)
        with torch.cuda.stream(aux):
            decode_raw(pix_logits, link_logits, dcfg, out, want_rects, ws)
        pixellink_loss_raw(pix_logits, link_logits, pix_lab, link_lab, lcfg, True, False, None, out)
        cur.wait_stream(aux)
        return out
    aux.wait_stream(cur)
    decode_raw(pix_logits, link_logits, dataclasses.replace(dcfg, phase=4, form="tiled"), out, want_rects, ws)
    aux.wait_stream(cur)
    pixellink_loss_raw(pix_logits, link_logits, pix_lab, link_lab, dataclasses.replace(lcfg, chain_pdl=True), True, False,
                       None, out)
    with torch.cuda.stream(aux):
        decode_raw(pix_logits, link_logits, dataclasses.replace(dcfg, phase=8, form="tiled"), out, want_rects, ws)
    cur.wait_stream(aux)
    return out


def contour_boxes_raw(mask: torch.Tensor, ratio_w: float = 1.0, ratio_h: float = 1.0, max_contours: int = 1024,
                      want_points: bool = False) -> dict:
    """plh_contour_boxes: mask uint8 [B,H,W] (CUDA) -> boxes / raw_boxes int32 [B,K,4,2], info int32 [B,K,6],
    n_contours int32 [B] (slots in arrival order), points int32 [B,2*H*W,2] if asked for."""
    lib = _lib.load()
    if mask.dtype != torch.uint8 or mask.dim() != 3:
        raise ValueError("mask must be uint8 [B,H,W]")
    dev = mask.device
    _require_gpu(dev)
    mask = mask.contiguous()
    B, H, W = mask.shape
    K = int(max_contours)
    out = {"boxes": torch.empty((B, K, 4, 2), dtype=torch.int32, device=dev),
           "raw_boxes": torch.empty((B, K, 4, 2), dtype=torch.int32, device=dev),
           "info": torch.empty((B, K, 6), dtype=torch.int32, device=dev),
           "n_contours": torch.empty((B,), dtype=torch.int32, device=dev),
           "points": torch.empty((B, 2 * H * W, 2), dtype=torch.int32, device=dev) if want_points else None}
    n = lib.plh_contour_workspace_bytes(B, H, W)
    if n == 0:
        raise ValueError("bad shape for the contour workspace: B=%d H=%d W=%d" % (B, H, W))
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.plh_contour_boxes(_p(mask), B, H, W, float(ratio_w), float(ratio_h), K, _p(out["boxes"]),
                                   _p(out["raw_boxes"]), _p(out["info"]), _p(out["n_contours"]), _p(out["points"]),
                                   _p(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "plh_contour_boxes")
    return out


def head_fuse_level_raw(feats, prev: Optional[torch.Tensor] = None, w_out: Optional[torch.Tensor] = None,
                        b_out: Optional[torch.Tensor] = None, flags_cfg: Optional["DecodeConfig"] = None,
                        logits: Optional[bool] = None):
    """plh_head_fuse_level: one level of the logit producer (nets/pixellink.py:56-67, nets/model.py:129-141).

    feats: one or two (x [B,H,W,K], w [K,18], scale [18] | None, shift [18] | None, relu) tuples (CUDA fp32);
    prev [B,H/2,W/2,18] | None; w_out [18,18] (in, out) + b_out [18] for the last level.
    logits (default: w_out is given): the output form — False: y18 [B,H,W,18] for the next level, True: (pixel
    logits [B,H,W,2], link logits [B,H,W,16]); with flags_cfg (a DecodeConfig, logits form only) also the decode's
    threshold words int16 [B,H,W] as third element."""
    lib = _lib.load()
    if not 1 <= len(feats) <= 2:
        raise ValueError("a level fuses one or two feature maps")
    x0 = feats[0][0]
    dev = x0.device
    _require_gpu(dev)
    B, H, W = x0.shape[:3]
    keep, flat = [], []
    for (x, w, scale, shift, relu) in feats:
        if x.dim() != 4 or tuple(x.shape[:3]) != (B, H, W) or tuple(w.shape) != (x.shape[3], 18):
            raise ValueError("feature [B,H,W,K] with weights [K,18] expected, got %s / %s" % (tuple(x.shape), tuple(w.shape)))
        t = [x.contiguous().float(), w.contiguous().float(), None if scale is None else scale.contiguous().float(),
             None if shift is None else shift.contiguous().float()]
        keep.append(t)
        flat += [_p(t[0]), int(x.shape[3]), _p(t[1]), _p(t[2]), _p(t[3]), int(bool(relu))]
    if len(feats) == 1:
        flat += [None, 0, None, None, None, 0]
    if prev is not None:
        if tuple(prev.shape) != (B, H // 2, W // 2, 18) or H % 2 or W % 2:
            raise ValueError("prev must be [B,H/2,W/2,18], got %s for a %dx%d level" % (tuple(prev.shape), H, W))
        prev = prev.contiguous()
    logits = (w_out is not None) if logits is None else bool(logits)
    if w_out is not None:
        w_out = w_out.contiguous().float()
        b_out = None if b_out is None else b_out.contiguous().float()
    if logits:
        pix = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev)
        link = torch.empty((B, H, W, 16), dtype=torch.float32, device=dev)
        y18 = None
    else:
        y18 = torch.empty((B, H, W, 18), dtype=torch.float32, device=dev)
        pix = link = None
    flags, dp = None, None
    if flags_cfg is not None:
        if not logits:
            raise ValueError("threshold words belong to the logits (the last level)")
        flags = torch.empty((B, H, W), dtype=torch.int16, device=dev)
        dp = flags_cfg.c_struct()
    with torch.cuda.device(dev):
        rc = lib.plh_head_fuse_level(*flat, _p(prev), _p(w_out), _p(b_out), B, H, W, _p(y18), _p(pix), _p(link),
                                     C.byref(dp) if dp is not None else None, _p(flags), _stream(dev))
    _lib.check(rc, "plh_head_fuse_level")
    if not logits:
        return y18
    return (pix, link) if flags is None else (pix, link, flags)


def fill_quads_raw(quads: torch.Tensor, counts, H: int, W: int, Ho: int, Wo: int, mode: int = 0, stride: int = 1,
                   zero_flags: Optional[torch.Tensor] = None, want=("last",)) -> dict:
    """plh_fill_quads: cv2.fillPoly of every image's quadrilaterals, in order, sampled on the output grid.

    quads int32 [sum n_b,4,2] (CUDA), counts: polygons per image (host sequence); mode 0: every stride-th canvas
    pixel, mode 1: cv2.resize(INTER_NEAREST) to (Ho, Wo).  want: any of "last" (always), "first", "ids_u8", "score",
    "training_mask" -> dict of [B,Ho,Wo] tensors."""
    lib = _lib.load()
    dev = quads.device
    _require_gpu(dev)
    if quads.dtype != torch.int32:
        raise ValueError("quads must be int32")
    counts = [int(c) for c in counts]
    B = len(counts)
    if B == 0 or quads.numel() != 8 * sum(counts):
        raise ValueError("counts do not add up to the polygon array")
    off = [0]
    for c in counts:
        off.append(off[-1] + c)
    quads = quads.contiguous() if quads.numel() else torch.zeros((1, 4, 2), dtype=torch.int32, device=dev)
    q_off = torch.tensor(off, dtype=torch.int32).to(dev, non_blocking=True)
    if zero_flags is not None:
        if zero_flags.dtype != torch.uint8 or zero_flags.numel() != off[-1]:
            raise ValueError("zero_flags must be uint8, one per polygon")
        zero_flags = zero_flags.contiguous() if off[-1] else None
    shape = (B, int(Ho), int(Wo))
    out = {"last": torch.empty(shape, dtype=torch.int32, device=dev),
           "first": torch.empty(shape, dtype=torch.int32, device=dev) if "first" in want else None,
           "ids_u8": torch.empty(shape, dtype=torch.uint8, device=dev) if "ids_u8" in want else None,
           "score": torch.empty(shape, dtype=torch.float32, device=dev) if "score" in want else None,
           "training_mask": torch.empty(shape, dtype=torch.uint8, device=dev) if "training_mask" in want else None}
    with torch.cuda.device(dev):
        rc = lib.plh_fill_quads(_p(quads), _p(q_off), _p(zero_flags), B, int(H), int(W), int(Ho), int(Wo), int(mode),
                                int(stride), _p(out["last"]), _p(out["first"]), _p(out["ids_u8"]), _p(out["score"]),
                                _p(out["training_mask"]), _stream(dev))
    _lib.check(rc, "plh_fill_quads")
    return out


def bboxes_matching_raw(dets: torch.Tensor, gts: torch.Tensor, det_counts, gt_counts, gignored: torch.Tensor,
                        matching_threshold: float = 0.5) -> dict:
    """plh_quad_jaccard + plh_bboxes_matching over a batch of images (tool/bboxes.py:158-282).

    dets int32 [sum D_b,4,2], gts int32 [sum G_b,4,2], gignored uint8 [sum G_b] (CUDA); det_counts / gt_counts:
    per-image counts (host sequences).  -> jaccard fp32 [sum D_b*G_b] (image b's D_b x G_b matrix at
    pair_off[b]), tp / fp uint8 [sum D_b], n_gbboxes int32 [B], pair_off (host list)."""
    lib = _lib.load()
    dev = dets.device
    _require_gpu(dev)
    if dets.dtype != torch.int32 or gts.dtype != torch.int32 or gignored.dtype != torch.uint8:
        raise ValueError("dets / gts must be int32, gignored uint8")
    det_counts = [int(v) for v in det_counts]
    gt_counts = [int(v) for v in gt_counts]
    B = len(det_counts)
    if B == 0 or len(gt_counts) != B:
        raise ValueError("det_counts and gt_counts must name the same (non-zero) number of images")
    if dets.numel() != 8 * sum(det_counts) or gts.numel() != 8 * sum(gt_counts) or gignored.numel() != sum(gt_counts):
        raise ValueError("counts do not add up to the box arrays")
    det_off, gt_off, pair_off = [0], [0], [0]
    for d, g in zip(det_counts, gt_counts):
        det_off.append(det_off[-1] + d), gt_off.append(gt_off[-1] + g), pair_off.append(pair_off[-1] + d * g)
    dets, gts, gignored = dets.contiguous(), gts.contiguous(), gignored.contiguous()
    d_off = torch.tensor(det_off, dtype=torch.int32).to(dev, non_blocking=True)
    g_off = torch.tensor(gt_off, dtype=torch.int32).to(dev, non_blocking=True)
    p_off = torch.tensor(pair_off, dtype=torch.int64).to(dev, non_blocking=True)
    out = {"jaccard": torch.empty((max(pair_off[-1], 1),), dtype=torch.float32, device=dev)[:pair_off[-1]],
           "tp": torch.empty((max(det_off[-1], 1),), dtype=torch.uint8, device=dev)[:det_off[-1]],
           "fp": torch.empty((max(det_off[-1], 1),), dtype=torch.uint8, device=dev)[:det_off[-1]],
           "n_gbboxes": torch.empty((B,), dtype=torch.int32, device=dev), "pair_off": pair_off,
           "det_off": det_off, "gt_off": gt_off}
    gmatch = torch.empty((max(gt_off[-1], 1),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.plh_quad_jaccard(_p(dets), _p(gts), _p(d_off), _p(g_off), _p(p_off), B, pair_off[-1], _p(out["jaccard"]),
                                  _stream(dev))
        _lib.check(rc, "plh_quad_jaccard")
        rc = lib.plh_bboxes_matching(_p(out["jaccard"]), _p(d_off), _p(g_off), _p(p_off), B, _p(gignored),
                                     float(matching_threshold), _p(gmatch), _p(out["tp"]), _p(out["fp"]),
                                     _p(out["n_gbboxes"]), _stream(dev))
        _lib.check(rc, "plh_bboxes_matching")
    return out


def link_labels_raw(ids: torch.Tensor, want_pixel: bool = True):
    """plh_link_labels: ids uint8 [B,H,W] (CUDA) -> (link_lab fp32 [B,H,W,8], pix_lab fp32 [B,H,W] or None)."""
    lib = _lib.load()
    if ids.dtype != torch.uint8 or ids.dim() != 3:
        raise ValueError("ids must be uint8 [B,H,W]")
    dev = ids.device
    _require_gpu(dev)
    ids = ids.contiguous()
    B, H, W = ids.shape
    link = torch.empty((B, H, W, 8), dtype=torch.float32, device=dev)
    pix = torch.empty((B, H, W), dtype=torch.float32, device=dev) if want_pixel else None
    with torch.cuda.device(dev):
        rc = lib.plh_link_labels(_p(ids), B, H, W, _p(link), _p(pix), _stream(dev))
    _lib.check(rc, "plh_link_labels")
    return link, pix


def link_labels_icdar_raw(last_ids: torch.Tensor, first_ids: torch.Tensor, stride: int = 1):
    """plh_link_labels_icdar: id maps int32 [B,H,W] (CUDA) -> (link fp32 [B,Ho,Wo,8], score fp32 [B,Ho,Wo])."""
    lib = _lib.load()
    if last_ids.dtype != torch.int32 or first_ids.dtype != torch.int32 or last_ids.dim() != 3 or last_ids.shape != first_ids.shape:
        raise ValueError("id maps must be int32 [B,H,W] of the same shape")
    dev = last_ids.device
    _require_gpu(dev)
    B, H, W = last_ids.shape
    if H != W:
        raise ValueError("datasets/icdar.py valid_link indexes out of range on non-square maps (quirk Q17)")
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    link = torch.empty((B, Ho, Wo, 8), dtype=torch.float32, device=dev)
    score = torch.empty((B, Ho, Wo), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.plh_link_labels_icdar(_p(last_ids.contiguous()), _p(first_ids.contiguous()), B, H, W, int(stride), _p(link),
                                       _p(score), _stream(dev))
    _lib.check(rc, "plh_link_labels_icdar")
    return link, score


def min_area_boxes_raw(pts: torch.Tensor, offsets: torch.Tensor, want_rects=True):
    """plh_min_area_boxes: pts int32 [total,2], offsets int32 [n+1] -> boxes int32 [n,4,2], rects [n,5]."""
    lib = _lib.load()
    dev = pts.device
    _require_gpu(dev)
    n = offsets.numel() - 1
    boxes = torch.empty((n, 4, 2), dtype=torch.int32, device=dev)
    rects = torch.empty((n, 5), dtype=torch.float32, device=dev) if want_rects else None
    with torch.cuda.device(dev):
        rc = lib.plh_min_area_boxes(_p(pts), _p(offsets), n, _p(boxes), _p(rects), _stream(dev))
    _lib.check(rc, "plh_min_area_boxes")
    return boxes, rects


def pixel_detect_raw(score, link, thr_p, thr_l):
    """plh_pixel_detect: score [H,W], link [8,H,W,2] probabilities -> uint8 [H,W]."""
    lib = _lib.load()
    H, W = score.shape
    dev = score.device
    _require_gpu(dev)
    outm = torch.empty((H, W), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.plh_pixel_detect(_p(score), _p(link), H, W, float(thr_p), float(thr_l), _p(outm), _stream(dev))
    _lib.check(rc, "plh_pixel_detect")
    return outm


# ----------------------------------------------------------------------------- dice / EAST
def dice_raw(y_true, y_pred, mask, want_grad=True):
    lib = _lib.load()
    dev = y_pred.device
    _require_gpu(dev)
    M = y_pred.numel()
    outv = torch.empty((4,), dtype=torch.float32, device=dev)
    grad = torch.empty_like(y_pred) if want_grad else None
    ws = _workspace(_lib.OP_DICE, 1, 1, 1, 0, dev)
    with torch.cuda.device(dev):
        rc = lib.plh_dice(_p(y_true), _p(y_pred), _p(mask), M, _p(outv), _p(grad), _p(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "plh_dice")
    return outv, grad


def dice_head_raw(t_pix, p_pix, t_link, p_link, mask, want_grad=True):
    lib = _lib.load()
    dev = p_pix.device
    _require_gpu(dev)
    M = p_pix.numel()
    if p_link.numel() != 8 * M or mask.numel() != M:
        raise ValueError("dice head expects pixel [M,1], link [M,8], mask [M]")
    outv = torch.empty((28,), dtype=torch.float32, device=dev)
    gp = torch.empty_like(p_pix) if want_grad else None
    gl = torch.empty_like(p_link) if want_grad else None
    ws = _workspace(_lib.OP_DICE, 1, 1, 1, 0, dev)
    with torch.cuda.device(dev):
        rc = lib.plh_dice_head(_p(t_pix), _p(p_pix), _p(t_link), _p(p_link), _p(mask), M, _p(outv), _p(gp), _p(gl),
                               _p(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "plh_dice_head")
    return outv, gp, gl


def east_loss_raw(score_gt, score_pred, geo_gt, geo_pred, mask, want_grad=True):
    lib = _lib.load()
    dev = score_pred.device
    _require_gpu(dev)
    M = score_pred.numel()
    if geo_pred.numel() != 5 * M:
        raise ValueError("EAST geometry must be [...,5]")
    outv = torch.empty((8,), dtype=torch.float32, device=dev)
    gs = torch.empty_like(score_pred) if want_grad else None
    gg = torch.empty_like(geo_pred) if want_grad else None
    ws = _workspace(_lib.OP_EAST_LOSS, 1, 1, 1, 0, dev)
    with torch.cuda.device(dev):
        rc = lib.plh_east_loss(_p(score_gt), _p(score_pred), _p(geo_gt), _p(geo_pred), _p(mask), M, _p(outv),
                               _p(gs), _p(gg), _p(ws), ws.numel(), _stream(dev))
    _lib.check(rc, "plh_east_loss")
    return outv, gs, gg


def lanms_raw(polys: torch.Tensor, offsets: torch.Tensor, thres: float = 0.3):
    """plh_lanms: polys fp64 [total,9], offsets int32 [B+1] -> (out fp64 [total,9], n_out int32 [B])."""
    lib = _lib.load()
    dev = polys.device
    _require_gpu(dev)
    total = polys.shape[0]
    Bn = offsets.numel() - 1
    out = torch.empty_like(polys)
    n_out = torch.zeros((Bn,), dtype=torch.int32, device=dev)
    ws = torch.empty(((total * 72 + 255) // 256) * 256 + total * 4 + 256, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.plh_lanms(_p(polys), _p(offsets), Bn, total, float(thres), _p(out), _p(n_out), _p(ws), ws.numel(),
                           _stream(dev))
    _lib.check(rc, "plh_lanms")
    return out, n_out


def restore_rectangle_raw(origin, geometry, want_index=False):
    """plh_restore_rectangle_ex: origin [N,2], geometry [N,5], each float32 or float64 (CUDA) -> [N,4,2] float64."""
    lib = _lib.load()
    dev = origin.device
    _require_gpu(dev)
    for t in (origin, geometry):
        if t.dtype not in (torch.float32, torch.float64):
            raise TypeError("restore_rectangle takes float32 / float64 tensors, got %s" % t.dtype)
    N = origin.shape[0]
    outp = torch.empty((N, 4, 2), dtype=torch.float64, device=dev)
    idx = torch.empty((N,), dtype=torch.int32, device=dev) if want_index else None
    if N == 0:
        return outp, idx
    ws = torch.empty((N // 1024 + 2) * 4 + 256, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.plh_restore_rectangle_ex(_p(origin), int(origin.dtype == torch.float64), _p(geometry),
                                          int(geometry.dtype == torch.float64), N, _p(outp), _p(idx), _p(ws),
                                          ws.numel(), _stream(dev))
    _lib.check(rc, "plh_restore_rectangle_ex")
    return outp, idx
