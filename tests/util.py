import numpy as np


def rel_err(a, b):
    """max |a-b| relative to the tensor's max magnitude (NaN positions must coincide)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN pattern differs: %d vs %d NaNs" % (na.sum(), nb.sum())
    if na.all():
        return 0.0
    a, b = a[~na], b[~nb]
    ia, ib = np.isinf(a), np.isinf(b)
    assert np.array_equal(ia, ib) and np.array_equal(a[ia], b[ib]), "inf pattern differs"
    a, b = a[~ia], b[~ib]
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


TOL = 1e-5  # north_star: losses / gradients within 1e-5 relative in fp32


def grad_close(a, b, w, rtol=TOL, ulps=4.0):
    """ELEMENTWISE gradient contract (include/plhead.h "Numerical contract"):
        |a - b| <= rtol * |b| + ulps * 2^-24 * |w|
    where w is the element's weight (gradient = w * (softmax - onehot)).  The absolute term is the rounding of
    `softmax - onehot` itself in fp32 — the reference (TF autodiff) and the oracle form that difference from a
    rounded softmax, so each of them carries up to ~1 ulp(1.0) * w of absolute error on near-saturated pixels;
    the kernel (sigmoid form) is more accurate there, and cannot agree better than the reference's own noise.
    a, b: [..., C, 2] or [..., 2]; w broadcast over the trailing class axis.  NaN patterns must coincide."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    w = np.abs(np.asarray(w, np.float64))
    if a.shape[-1] == 2 and w.shape == a.shape[:-1]:
        w = w[..., None]
    elif a.ndim == w.ndim and a.shape[-1] == 2 * w.shape[-1]:
        w = np.repeat(w, 2, axis=-1)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN pattern differs"
    ok = ~na & np.isfinite(b)
    lim = rtol * np.abs(b) + ulps * 2.0 ** -24 * np.where(np.isfinite(w), w, 0.0)
    bad = ok & (np.abs(a - b) > lim)
    assert not bad.any(), "%d elements outside the elementwise tolerance, worst |a-b|/lim = %.3g" % (
        bad.sum(), np.max((np.abs(a - b) / np.maximum(lim, 1e-300))[ok]))
    return True


def link_graph_case(golden_dir, tag):
    """One case of tests/golden/link_graph.npz (reference script executed by line range): flags, the logits a
    decode call needs to reproduce them (+-3 margins: scores 0.9975 / 0.0025), and the reference's result."""
    g = np.load(golden_dir + "/link_graph.npz")
    H, W = int(g[tag + "_H"]), int(g[tag + "_W"])
    P = np.unpackbits(g[tag + "_P"])[: H * W].reshape(H, W).astype(bool)
    L = np.unpackbits(g[tag + "_L"])[: H * W * 8].reshape(H, W, 8).astype(bool)
    pix = np.zeros((H, W, 2), np.float32)
    pix[..., 1] = np.where(P, 3.0, -3.0)
    link = np.zeros((H, W, 8, 2), np.float32)
    link[..., 1] = np.where(L, 3.0, -3.0)
    return dict(H=H, W=W, P=P, L=L, pix_logits=pix, link_logits=link.reshape(H, W, 16), labels=g[tag + "_labels"],
                boxes=g[tag + "_boxes"], min_size=int(g[tag + "_min_size"]),
                scale=tuple(float(v) for v in g[tag + "_scale"]), n_groups=int(g[tag + "_n_groups"]),
                res_txt=bytes(g["fast_res_txt"]) if tag == "fast" else None)
