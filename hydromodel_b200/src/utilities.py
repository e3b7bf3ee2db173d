"""Small numerical helpers shared by the host-side objects.

``logN_rnd`` and ``find_wtd`` keep the names, argument meaning and results of
the reference's Numba functions (code/src/utilities.py:22-54 and :56-99); on
the device they are inlined in the right-hand-side kernel
(hydromodel_b200/csrc/hb_device.cuh).  ``lerp_table`` reproduces the arithmetic
of SciPy's linear ``interp1d`` evaluation that the reference uses for every
static profile lookup (porosity.py:184, vrettas_fung.py:34,47, tree_roots.py:135),
so that host-precomputed per-cell tables equal the reference's per-call values.
"""
import numpy as np


def logN_rnd(mx, vx, en):
    """Log-normal variate with mean ``mx`` and variance ``vx`` of the
    underlying field, driven by the standard-normal draw ``en``."""
    mx, vx, en = np.atleast_1d(mx, vx, en)
    mx = np.where(mx == 0.0, 1.0e-7, mx)
    m2 = mx * mx
    return np.exp(np.log(m2 / np.sqrt(vx + m2)) + np.sqrt(np.log(vx / m2 + 1.0)) * en)


def find_wtd(y=None):
    """Index just below the deepest unsaturated cell of the boolean vector
    ``y`` (True = saturated), clipped to ``len(y) - 1``; 0 if all saturated."""
    if y is None:
        return None
    y = np.atleast_1d(y)
    dry = np.flatnonzero(~y.astype(bool))
    top = int(dry[-1]) + 1 if dry.size else 0
    return int(min(top, y.shape[0] - 1))


def lerp_table(x_knots, y_knots, x_new):
    """Piecewise-linear lookup on increasing knots with the arithmetic of
    SciPy's ``interp1d`` (scipy/interpolate/_interpolate.py:487-518): a 1-D table
    goes through ``slope * (x - x_lo) + y_lo`` (np.interp), a 2-D table
    ``[n_series, n_knots]`` through the two-weight form
    ``w * y_hi + (1 - w') * y_lo``.  Raises ``ValueError`` outside the knot range
    (as ``interp1d`` does by default)."""
    x_knots = np.asarray(x_knots, dtype=float)
    y_knots = np.asarray(y_knots, dtype=float)
    x_new = np.asarray(x_new, dtype=float)
    if np.any(x_new < x_knots[0]) or np.any(x_new > x_knots[-1]):
        raise ValueError("A value in x_new is outside the interpolation range.")
    if y_knots.ndim == 1:
        return np.interp(x_new, x_knots, y_knots)
    hi = np.clip(np.searchsorted(x_knots, x_new), 1, x_knots.size - 1)
    lo = hi - 1
    x_lo, x_hi = x_knots[lo], x_knots[hi]
    return ((x_new - x_lo) / (x_hi - x_lo)) * y_knots[..., hi] + \
        ((x_hi - x_new) / (x_hi - x_lo)) * y_knots[..., lo]
