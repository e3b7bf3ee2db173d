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
    """Piecewise-linear lookup ``slope * (x - x_lo) + y_lo`` on increasing
    knots.  ``y_knots`` may be 1-D or 2-D ``[n_series, n_knots]``.  Raises
    ``ValueError`` outside the knot range (as ``interp1d`` does by default)."""
    x_knots = np.asarray(x_knots, dtype=float)
    y_knots = np.asarray(y_knots, dtype=float)
    x_new = np.asarray(x_new, dtype=float)
    if np.any(x_new < x_knots[0]) or np.any(x_new > x_knots[-1]):
        raise ValueError("A value in x_new is outside the interpolation range.")
    hi = np.clip(np.searchsorted(x_knots, x_new), 1, x_knots.size - 1)
    lo = hi - 1
    slope = (y_knots[..., hi] - y_knots[..., lo]) / (x_knots[hi] - x_knots[lo])
    return slope * (x_new - x_knots[lo]) + y_knots[..., lo]
