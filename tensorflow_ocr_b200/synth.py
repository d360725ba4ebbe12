"""Deterministic synthetic head inputs (SURVEY.md §8d "Synthetic inputs").

The reference feeds the head from a trained backbone and from
``datasets/icdar.py:486-539`` / ``tool/pixellink_fn.py:49-111`` label generation.
Neither is available offline, so tests and ``bench.py`` use seeded synthetic maps
of the same shape/dtype (NHWC fp32, labels 0.0/1.0):

* labels: random rotated rectangles rasterised with ``cv2.fillPoly`` into an id
  map; ``pix_lab = id > 0``; ``link_lab[d] = 1`` iff the pixel is text and its
  d-neighbour carries the same id, border text pixels link in every direction
  (tool/pixellink_fn.py:9-47 ``valid_link`` semantics);
* logits family ``G`` (parity, tie-robust): margins on a 1/64 grid, so every
  ``expf`` implementation orders the scores identically and OHEM ties are common;
* family ``C`` (throughput): same without quantisation;
* family ``S``: link logits symmetrised (``z_d[v] = z_opp(d)[u]``), which makes the
  reference's literal directed DFS equal to weakly-connected components.

Host-side numpy only; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

# Neighbour table of the decode (tool/pixellink_fn.py:93-108;
# test_pixellink_fast.py:124-146): channel d -> (dy, dx).
NEIGHBOURS = (
    (0, -1),   # 0 left
    (1, -1),   # 1 left_down
    (-1, -1),  # 2 left_up
    (0, 1),    # 3 right
    (1, 1),    # 4 right_down
    (-1, 1),   # 5 right_up
    (-1, 0),   # 6 up
    (1, 0),    # 7 down
)
OPPOSITE = (3, 5, 4, 0, 2, 1, 7, 6)

SEED_BASE = 20260000


def image_seed(config: int, image: int) -> int:
    return SEED_BASE + 1000 * int(config) + int(image)


def _rng(config: int, image: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(image_seed(config, image)))


def make_id_map(rng: np.random.Generator, H: int, W: int, kind: str = "normal") -> np.ndarray:
    """int32 [H, W] instance-id map (0 = background)."""
    import cv2

    ids = np.zeros((H, W), np.int32)
    if kind == "no_pos":
        return ids
    if kind == "all_pos":
        ids[:] = 1
        return ids
    if kind == "many_pos":  # positives > negatives/3 -> k clamps to #neg
        ids[:] = 0
        ids[H // 8: H - H // 8, W // 8: W - W // 8] = 1
        ids[H // 2 - 1: H // 2 + 1, :] = 0  # two blobs
        ids[H // 2 + 1: H - H // 8, W // 8: W - W // 8] = 2
        return ids
    scale = min(H, W) / 128.0
    K = int(rng.integers(3, 13))
    for k in range(K):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        long_side = rng.uniform(8.0, 60.0) * scale
        aspect = rng.uniform(2.0, 8.0)
        short = long_side / aspect
        ang = np.deg2rad(rng.uniform(-45.0, 45.0))
        c, s = np.cos(ang), np.sin(ang)
        hx, hy = long_side / 2.0, max(short / 2.0, 0.75)
        corners = np.array([[-hx, -hy], [hx, -hy], [hx, hy], [-hx, hy]])
        rot = np.array([[c, -s], [s, c]])
        pts = corners @ rot.T + np.array([cx, cy])
        cv2.fillPoly(ids, [np.round(pts).astype(np.int32)], int(k + 1))
    return ids


def labels_from_ids(ids: np.ndarray):
    """pix_lab [H,W,1], link_lab [H,W,8] fp32 from an id map."""
    H, W = ids.shape
    pix = (ids > 0)
    link = np.zeros((H, W, 8), np.float32)
    border = np.zeros((H, W), bool)
    border[0, :] = border[-1, :] = True
    border[:, 0] = border[:, -1] = True
    pad = np.pad(ids, 1, constant_values=-1)
    for d, (dy, dx) in enumerate(NEIGHBOURS):
        nb = pad[1 + dy: 1 + dy + H, 1 + dx: 1 + dx + W]
        link[:, :, d] = (pix & (border | (nb == ids))).astype(np.float32)
    return pix.astype(np.float32)[:, :, None], link


def _margins(rng, lab, quantise: bool):
    z = (2.0 * lab - 1.0) * 2.0 + 1.5 * rng.standard_normal(lab.shape)
    c = rng.standard_normal(lab.shape)
    if quantise:
        z = np.round(z * 64.0) / 64.0
        c = np.round(c * 64.0) / 64.0
    return z, c


def _logits_from_margin(z, c):
    out = np.empty(z.shape + (2,), np.float32)
    out[..., 1] = (z / 2.0 + c).astype(np.float32)
    out[..., 0] = (-z / 2.0 + c).astype(np.float32)
    return out


def symmetrise_links(z: np.ndarray) -> np.ndarray:
    """z [H,W,8] -> z'[v,d] = z'[u,opp(d)] for every in-map edge (family S)."""
    H, W, _ = z.shape
    out = z.copy()
    for d in (3, 4, 5, 7):  # each undirected edge once: right, right_down, right_up, down
        dy, dx = NEIGHBOURS[d]
        o = OPPOSITE[d]
        ys = slice(max(0, -dy), H - max(0, dy))
        xs = slice(max(0, -dx), W - max(0, dx))
        ys_u = slice(max(0, -dy) + dy, H - max(0, dy) + dy)
        xs_u = slice(max(0, -dx) + dx, W - max(0, dx) + dx)
        out[ys_u, xs_u, o] = out[ys, xs, d]
    return out


def make_image(config: int, image: int, H: int, W: int, family: str = "G", kind: str = "normal"):
    """One image's head inputs.

    Returns dict with pix_logits [H,W,2], link_logits [H,W,16], pix_lab [H,W,1],
    link_lab [H,W,8] (fp32) and ids [H,W] (int32).
    """
    rng = _rng(config, image)
    ids = make_id_map(rng, H, W, kind)
    pix_lab, link_lab = labels_from_ids(ids)
    quant = family in ("G", "S")
    zp, cp = _margins(rng, pix_lab[..., 0], quant)
    zl, cl = _margins(rng, link_lab, quant)
    if family == "S":
        zl = symmetrise_links(zl)
    pix_logits = _logits_from_margin(zp, cp)
    link_logits = _logits_from_margin(zl, cl).reshape(H, W, 16)
    return dict(pix_logits=pix_logits, link_logits=link_logits,
                pix_lab=pix_lab, link_lab=link_lab, ids=ids)


EDGE_KINDS = ("no_pos", "many_pos", "all_pos")


def make_batch(config: int, B: int, H: int, W: int, family: str = "G",
               edge_images: bool = False, first_image: int = 0):
    """Batch of head inputs, NHWC fp32.

    ``edge_images=True`` replaces the last three images by the mandatory edge
    cases (no positives / positives > negatives/3 / all positive).
    """
    kinds = ["normal"] * B
    if edge_images:
        for j, k in enumerate(EDGE_KINDS):
            if B - 1 - j >= 0:
                kinds[B - 1 - j] = k
    imgs = [make_image(config, first_image + i, H, W, family, kinds[i]) for i in range(B)]
    out = {k: np.ascontiguousarray(np.stack([im[k] for im in imgs])) for k in imgs[0]}
    out["training_mask"] = np.ones((B, H, W, 1), np.float32)
    return out


def make_east_batch(config: int, B: int, H: int, W: int, first_image: int = 0, consistent: bool = False):
    """EAST RBOX head inputs (config 4): score [B,H,W,1] prob, geo [B,H,W,5]
    (4 distances + angle) for prediction and ground truth, plus training mask.

    consistent=True: the geometry of a text pixel is the distances (top, right, bottom, left, in input pixels =
    4 x map pixels) to the sides of ITS rotated rectangle and the rectangle's angle, as the EAST label generator
    produces them, and the prediction is that plus a few percent of noise — so that the pixels of one instance
    restore to nearly the same quadrilateral and locality-aware NMS has its usual work.  Otherwise uniform noise."""
    import cv2

    outs = dict(score_gt=[], score_pred=[], geo_gt=[], geo_pred=[], training_mask=[])
    for i in range(B):
        rng = _rng(config, first_image + i)
        if consistent:
            ids = np.zeros((H, W), np.int32)
            d_gt = np.ones((H, W, 4), np.float32)
            th_gt = np.zeros((H, W, 1), np.float32)
            yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
            scale = min(H, W) / 128.0
            for k in range(int(rng.integers(3, 13))):
                cx, cy = rng.uniform(0, W), rng.uniform(0, H)
                long_side = rng.uniform(8.0, 60.0) * scale
                hx, hy = long_side / 2.0, max(long_side / rng.uniform(2.0, 8.0) / 2.0, 0.75)
                ang = np.deg2rad(rng.uniform(-30.0, 30.0))
                c, s_ = np.cos(ang), np.sin(ang)
                pts = np.array([[-hx, -hy], [hx, -hy], [hx, hy], [-hx, hy]]) @ np.array([[c, -s_], [s_, c]]).T + (cx, cy)
                m = np.zeros((H, W), np.uint8)
                cv2.fillPoly(m, [np.round(pts).astype(np.int32)], 1)
                m = m > 0
                ids[m] = k + 1
                u = (xx - cx) * c + (yy - cy) * s_          # along the rectangle's width / height
                v = -(xx - cx) * s_ + (yy - cy) * c
                d = np.stack([hy + v, hx - u, hy - v, hx + u], -1) * 4.0     # top, right, bottom, left
                d_gt[m] = np.maximum(d[m], 0.5)
                th_gt[m] = -ang                                # image y points down
        else:
            ids = make_id_map(rng, H, W)
            d_gt = rng.uniform(1.0, 40.0, (H, W, 4)).astype(np.float32)
            th_gt = rng.uniform(-np.pi / 4, np.pi / 4, (H, W, 1)).astype(np.float32)
        gt = (ids > 0).astype(np.float32)[..., None]
        z = (2.0 * gt - 1.0) * 2.0 + 1.5 * rng.standard_normal(gt.shape)
        pred = (1.0 / (1.0 + np.exp(-z))).astype(np.float32)
        if consistent:
            d_pr = (d_gt * rng.uniform(0.97, 1.03, (H, W, 4))).astype(np.float32)
            th_pr = (th_gt + 0.02 * rng.standard_normal((H, W, 1))).astype(np.float32)
        else:
            d_pr = (d_gt * rng.uniform(0.6, 1.4, (H, W, 4))).astype(np.float32)
            th_pr = (th_gt + 0.2 * rng.standard_normal((H, W, 1))).astype(np.float32)
        tm = (rng.uniform(size=(H, W, 1)) > 0.05).astype(np.float32)
        outs["score_gt"].append(gt)
        outs["score_pred"].append(pred)
        outs["geo_gt"].append(np.concatenate([d_gt, th_gt], -1).astype(np.float32))
        outs["geo_pred"].append(np.concatenate([d_pr, th_pr], -1).astype(np.float32))
        outs["training_mask"].append(tm)
    return {k: np.ascontiguousarray(np.stack(v)) for k, v in outs.items()}
