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
