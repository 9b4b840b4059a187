"""Shared test helpers: host-side restatement of the frame -> design step for the small fixtures."""
import numpy as np

SMALL = 1e-6


def _den(b, small, cancel_floor):
    nan_b = np.isnan(b)
    absb = np.abs(np.where(nan_b, 0.0, b))
    colscale = absb.max(axis=0, keepdims=True) if b.ndim >= 2 else absb.max()
    # a column that is rounding noise throughout (every entry < 1e-6 of the array's largest value, e.g. the identically
    # zero "explained" term of the intercept) is held to the array's scale instead of its own
    colscale = np.broadcast_to(np.maximum(colscale, small * absb.max()), b.shape)
    den = np.where(absb >= small, absb, np.maximum(colscale, np.finfo(float).tiny))
    if cancel_floor:
        rowscale = absb.max(axis=-1, keepdims=True) if b.ndim >= 2 else absb.max()
        den = np.maximum(den, cancel_floor * rowscale)
    return den, colscale


def relerr(a, b, small=SMALL, cancel_floor=0.0):
    """north_star's "within a relative 1e-10", read ELEMENT BY ELEMENT: max_ij |a_ij - b_ij| / den_ij with
    den_ij = |b_ij| wherever |b_ij| >= small (1e-6), and the scale of the element's column, max_i |b_ij| (1-D arrays: of
    the whole vector), only where the reference value itself is smaller than that -- an absolute floor is needed there
    because such entries are differences of O(scale) terms (tolerance 1e-10 => floor 1e-10 x column scale).  NaN patterns must coincide.

    cancel_floor (default 0 = off; used only for the reference's 5-rows-per-group fixtures): den_ij is at least
    cancel_floor x the largest statistic of the same replicate (row).  On such resamples a detailed term like
    xbar_j (beta_a_j - beta*_j) can be 1e-6 of its own operands (two coefficients that agree to six digits), so no
    fp64 evaluation carries ten digits of it; the operands' scale is the honest reference there."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN pattern differs"
    if a.size == 0 or nan_a.all():
        return 0.0
    den, _ = _den(b, small, cancel_floor)
    err = np.abs(np.where(nan_a, 0.0, a) - np.where(nan_b, 0.0, b)) / den
    return float(err.max())


def worst(a, b, small=SMALL, cancel_floor=0.0):
    """Where relerr(a, b) comes from, for assertion messages: (error, index, a value, b value, column scale)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.size == 0 or np.isnan(b).all():
        return None
    den, colscale = _den(b, small, cancel_floor)
    err = np.abs(np.nan_to_num(a) - np.nan_to_num(b)) / den
    i = np.unravel_index(int(np.argmax(err)), err.shape)
    return dict(err=float(err[i]), index=tuple(int(x) for x in i), got=float(a[i]), want=float(b[i]), col_scale=float(colscale[i]))


def relerr_to_scale(a, b, scale):
    """|a - b| / scale for quantities that ARE differences of O(scale) numbers (residuals y - x.beta: scale = max |y|)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0



def fixture_design(fix, weighted=False):
    c = fix["columns"]
    g = np.array(c[fix["group"]])
    refg = fix["reference_group"]
    other = sorted(set(g) - {refg})[0]          # builder.rs:73-83
    A, B = g == other, g == refg
    cols = [np.ones(len(g))] + [np.array(c[p], float) for p in fix["predictors"]]
    norm = []
    for cat in fix.get("categorical", []):
        v = np.array(c[cat])
        levels = sorted(set(v))                   # builder.rs:384-388
        start = len(cols)
        for lv in levels[1:]:                     # builder.rs:402
            cols.append((v == lv).astype(float))
        if cat in fix.get("normalize", []):
            norm.append((len(levels), list(range(start, len(cols)))))
    X = np.stack(cols, 1)
    y = np.array(c[fix["outcome"]], float)
    w = np.array(c["weight"], float) if weighted else None
    return (X[A], y[A], None if w is None else w[A], X[B], y[B], None if w is None else w[B],
            norm, len(fix["predictors"]))


