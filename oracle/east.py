"""Oracle: EAST RBOX pieces (numpy restatement).

TEST INFRASTRUCTURE — see oracle/__init__.py.

* ``restore_rectangle_rbox`` follows datasets/icdar.py:410-483 (E1) and is pinned by
  tests/golden/restore_rectangle.npz (produced by running the reference function).
* ``east_loss`` (E2) and ``nms_locality`` (E3) DO NOT EXIST in the reference
  (SURVEY.md §8a): they are restated from upstream argman/EAST (``model.loss``,
  ``locality_aware_nms.py``) and are PARITY UNPINNED.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------- E1
def restore_rectangle_rbox(origin, geometry):
    """datasets/icdar.py:410-479.  origin [N,2], geometry [N,5] = (top,right,bottom,left,theta).

    dtype behaviour of the reference is kept: distance sums and cos/sin are computed in
    the input dtype (fp32), the local points live in a float64 array (np.zeros), so the
    rotation products and the translation are float64.  Rows: theta >= 0 first, then
    theta < 0 (icdar.py:479, quirk Q16).
    """
    origin = np.asarray(origin)
    geometry = np.asarray(geometry)
    d = geometry[:, :4]
    angle = geometry[:, 4]
    parts = []
    for nonneg in (True, False):
        sel = (angle >= 0) if nonneg else (angle < 0)
        o, dd, a = origin[sel], d[sel], angle[sel]
        n = o.shape[0]
        if n == 0:
            parts.append(np.zeros((0, 4, 2)))
            continue
        zero = np.zeros(n)                          # float64, promotes the stacks below
        hh = -dd[:, 0] - dd[:, 2]
        if nonneg:                                   # icdar.py:417-443
            ww = dd[:, 1] + dd[:, 3]
            px = np.stack([zero, ww, ww, zero, dd[:, 3]], 1)
            py = np.stack([hh, hh, zero, zero, -dd[:, 2]], 1)
            c, s = np.cos(a), np.sin(a)
            rx = c[:, None] * px + s[:, None] * py
            ry = (-s)[:, None] * px + c[:, None] * py
        else:                                        # icdar.py:450-476
            ww = -dd[:, 1] - dd[:, 3]
            px = np.stack([ww, zero, zero, ww, -dd[:, 1]], 1)
            py = np.stack([hh, hh, zero, zero, -dd[:, 2]], 1)
            c, s = np.cos(-a), np.sin(-a)
            rx = c[:, None] * px + (-s)[:, None] * py
            ry = s[:, None] * px + c[:, None] * py
        shift_x = o[:, 0] - rx[:, 4]
        shift_y = o[:, 1] - ry[:, 4]
        quad = np.stack([rx[:, :4] + shift_x[:, None], ry[:, :4] + shift_y[:, None]], -1)   # [n,4,2]
        parts.append(quad)
    return np.concatenate(parts)


def restore_rectangle(origin, geometry):
    """datasets/icdar.py:482-483."""
    return restore_rectangle_rbox(origin, geometry)


# --------------------------------------------------------------------------- E2 (parity unpinned)
def east_loss(score_gt, score_pred, geo_gt, geo_pred, training_mask):
    """Upstream argman/EAST ``model.loss``:
    L = mean(y_true * mask * (L_AABB + 20 L_theta)) + 0.01 * dice(y_true, y_pred, mask),
    L_AABB = -log((A_i + 1) / (A_u + 1)), L_theta = 1 - cos(theta_pred - theta_gt).
    Returns dict with loss and gradients wrt score_pred / geo_pred (upstream gradient 1).
    """
    t = np.asarray(score_gt, f32)
    p = np.asarray(score_pred, f32)
    g = np.asarray(geo_gt, f32)
    q = np.asarray(geo_pred, f32)
    m = np.asarray(training_mask, f32)
    I = f32(np.sum((t * p * m).astype(np.float64)))
    U = f32(f32(np.sum((t * m).astype(np.float64))) + f32(np.sum((p * m).astype(np.float64))) + f32(1e-5))
    dice = f32(f32(1) - f32(2) * I / U)
    area_gt = (g[..., 0] + g[..., 2]) * (g[..., 1] + g[..., 3])
    area_pr = (q[..., 0] + q[..., 2]) * (q[..., 1] + q[..., 3])
    w_union = np.minimum(g[..., 1], q[..., 1]) + np.minimum(g[..., 3], q[..., 3])
    h_union = np.minimum(g[..., 0], q[..., 0]) + np.minimum(g[..., 2], q[..., 2])
    ai = (w_union * h_union).astype(f32)
    au = (area_gt + area_pr - ai).astype(f32)
    l_aabb = (-np.log((ai + f32(1)) / (au + f32(1)))).astype(f32)
    l_theta = (f32(1) - np.cos(q[..., 4] - g[..., 4])).astype(f32)
    w = (t[..., 0] * m[..., 0]).astype(f32)
    M = t.size
    s_aabb = float(np.sum((l_aabb * w).astype(np.float64)))
    s_theta = float(np.sum((l_theta * w).astype(np.float64)))
    lg = f32((s_aabb + 20.0 * s_theta) / M)
    loss = f32(lg + f32(0.01) * dice)
    grad_score = (f32(0.01) * f32(-2) * m * (t * U - I) / (U * U)).astype(f32)
    wm = (w / f32(M)).astype(f32)
    ri = (f32(1) / (ai + f32(1))).astype(f32)
    ru = (f32(1) / (au + f32(1))).astype(f32)
    gg = np.zeros_like(q)
    for c in range(4):
        other = w_union if c in (0, 2) else h_union       # dA_i/dd_c = [pred < gt] * (the other extent)
        dai = np.where(q[..., c] < g[..., c], other, f32(0)).astype(f32)
        dap = (q[..., 1] + q[..., 3]) if c in (0, 2) else (q[..., 0] + q[..., 2])
        gg[..., c] = wm * (-ri * dai + ru * (dap - dai))
    gg[..., 4] = wm * f32(20) * np.sin(q[..., 4] - g[..., 4])
    return dict(loss=loss, dice=dice, lg=lg, I=I, U=U, grad_score=grad_score, grad_geo=gg.astype(f32))


# --------------------------------------------------------------------------- E3 (parity unpinned)
def _poly_area(p):
    x, y = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _clip(subject, a, b):
    """Sutherland-Hodgman: keep the part of `subject` on the left of the directed edge a->b."""
    out = []
    n = len(subject)
    for i in range(n):
        cur, nxt = subject[i], subject[(i + 1) % n]
        sc = (b[0] - a[0]) * (cur[1] - a[1]) - (b[1] - a[1]) * (cur[0] - a[0])
        sn = (b[0] - a[0]) * (nxt[1] - a[1]) - (b[1] - a[1]) * (nxt[0] - a[0])
        if sc >= 0:
            out.append(cur)
        if (sc >= 0) != (sn >= 0):
            tt = sc / (sc - sn)
            out.append((cur[0] + tt * (nxt[0] - cur[0]), cur[1] + tt * (nxt[1] - cur[1])))
    return out


def quad_iou(g, p):
    """IoU of two convex quadrilaterals given as 8 coordinates (upstream uses shapely)."""
    G = np.asarray(g[:8], np.float64).reshape(4, 2)
    P = np.asarray(p[:8], np.float64).reshape(4, 2)
    ag, ap = _poly_area(G), _poly_area(P)
    if ag < 0:
        G, ag = G[::-1], -ag
    if ap < 0:
        P, ap = P[::-1], -ap
    if ag == 0 or ap == 0:
        return 0.0
    poly = [tuple(v) for v in G]
    for i in range(4):
        if not poly:
            break
        poly = _clip(poly, P[i], P[(i + 1) % 4])
    inter = abs(_poly_area(np.asarray(poly))) if len(poly) >= 3 else 0.0
    union = ag + ap - inter
    return inter / union if union != 0 else 0.0


def weighted_merge(g, p):
    g = np.array(g, np.float64)
    p = np.asarray(p, np.float64)
    g[:8] = (g[8] * g[:8] + p[8] * p[:8]) / (g[8] + p[8])
    g[8] = g[8] + p[8]
    return g


def standard_nms(S, thres):
    order = np.argsort(S[:, 8])[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        ovr = np.array([quad_iou(S[i], S[t]) for t in order[1:]])
        inds = np.where(ovr <= thres)[0]
        order = order[inds + 1]
    return S[keep]


def nms_locality(polys, thres=0.3):
    """Upstream EAST locality_aware_nms.nms_locality: row-major fold merging consecutive
    overlapping boxes by score-weighted average, then standard NMS.  polys [N,9]."""
    S = []
    p = None
    for g in np.asarray(polys, np.float64):
        if p is not None and quad_iou(g, p) > thres:
            p = weighted_merge(g, p)
        else:
            if p is not None:
                S.append(p)
            p = g
    if p is not None:
        S.append(p)
    if len(S) == 0:
        return np.zeros((0, 9))
    return standard_nms(np.array(S), thres)
