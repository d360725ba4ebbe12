"""CPU: size-independent properties of the oracle itself (hypothesis), so that the checker the GPU
parity tests lean on is exercised well beyond the handful of golden vectors."""
import numpy as np
from hypothesis import given, settings, strategies as st


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 400), st.integers(0, 5))
def test_ohnm_selects_exactly_the_k_lowest_scores_plus_ties(seed, n, levels):
    """nets/model.py:161-184: selected = negatives with score <= the k-th smallest negative score,
    k = min(3 n_pos, n_neg); so |selected| >= k, and == k without ties at the threshold."""
    from oracle import pixellink_loss as O
    rng = np.random.default_rng(seed)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    if levels:
        scores = (np.round(scores * levels) / levels).astype(np.float32)   # force ties
    lab = rng.integers(0, 2, n)
    pos, neg = lab == 1, lab == 0
    n_pos, n_neg = int(pos.sum()), int(neg.sum())
    sel, _ = O.OHNM_single_image(scores, n_pos, neg.astype(np.float32))
    sel = np.asarray(sel).astype(bool)
    assert not (sel & ~neg).any()
    k = min(3 * n_pos, n_neg)
    if n_pos == 0 or k == 0:
        assert not sel.any()
        return
    thr = np.sort(scores[neg])[k - 1]
    assert np.array_equal(sel, neg & (scores <= thr))
    assert sel.sum() >= k and (sel.sum() == k or (scores[neg] == thr).sum() > 1)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(3, 24), st.integers(3, 24), st.integers(0, 6))
def test_link_components_canonical_labels(seed, H, W, min_size):
    """Decode oracle (canonical contract): every kept component is labelled with its minimum
    pixel index, has more than min_size pixels, roots come ascending with matching sizes, labelled pixels
    are positive pixels, and every reference edge joins two pixels of the same component."""
    from oracle import decode as D
    rng = np.random.default_rng(seed)
    P = rng.uniform(size=(H, W)) < 0.55
    L = rng.uniform(size=(H, W, 8)) < 0.6
    lab, roots, sizes = D.link_components(P, L, min_size=min_size)[:3]
    assert not (lab[~P] >= 0).any()
    flat = lab.reshape(-1)
    assert list(roots) == sorted(set(flat[flat >= 0].tolist()))
    for r, sz in zip(roots, sizes):
        idx = np.flatnonzero(flat == r)
        assert idx.min() == r and len(idx) == sz and sz > min_size
    full = D.link_components(P, L, min_size=0)[0].reshape(-1)
    src, dst, _ = D.edge_list(P, L)
    assert np.array_equal(full[src], full[dst])


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40))
def test_min_area_box_contains_all_points(seed, n):
    """oracle/minarearect.py: the integer box of cv2-style minAreaRect encloses the points up to the
    truncation of its corners (1 px)."""
    import cv2
    from oracle import minarearect as M
    rng = np.random.default_rng(seed)
    pts = rng.integers(0, 200, size=(n, 2)).astype(np.int32)
    box = np.asarray(M.min_area_box_int(pts)).reshape(4, 2)
    ref = np.asarray(cv2.boxPoints(cv2.minAreaRect(pts.astype(np.float32)))).astype(np.int32)  # same truncation
    assert np.array_equal(box, ref)
