"""Soil hydraulic parameters (value object).

Same public surface as the reference's ``SoilProperties``
(code/src/soil_properties.py:3-204): attributes ``n, m, alpha, psi_sat,
epsilon`` with range-checked setters that raise ``ValueError``.  On the device
these four numbers become entries of the per-column parameter block
(include/hb_layout.h: HB_P_N, HB_P_M, HB_P_ALPHA, HB_P_PSI_SAT, HB_P_EPS).
"""
import numpy as np


def _checked(name, ok, hint):
    """Build a property whose setter rejects values failing ``ok``."""
    slot = "_" + name

    def getter(self):
        return getattr(self, slot)

    def setter(self, value):
        if not ok(value):
            raise ValueError(f" {type(self).__name__}: Soil property '{name}': "
                             f"{value}, should be {hint}.")
        setattr(self, slot, value)

    return property(getter, setter, doc=f"soil property '{name}' ({hint})")


class SoilProperties:
    """n > 1, alpha > 0, psi_sat <= 0, epsilon > 0 (lengths in cm)."""

    __slots__ = ("_n", "_alpha", "_psi_sat", "_epsilon")

    n = _checked("n", lambda v: v > 1.0, "> 1")
    alpha = _checked("alpha", lambda v: v > 0.0, "strictly positive")
    psi_sat = _checked("psi_sat", lambda v: v <= 0.0, "<= 0")
    epsilon = _checked("epsilon", lambda v: v > 0.0, "(tiny) positive")

    def __init__(self, n: float = 2.0, alpha: float = 0.009,
                 psi_sat: float = -100.0, epsilon: float = 1.0e-7):
        self.n = n
        self.alpha = alpha
        # The constructor clamps instead of raising for these two
        # (reference soil_properties.py:45-49).
        self._psi_sat = np.minimum(psi_sat, 0.0)
        self._epsilon = np.maximum(epsilon, 1.0e-8)

    @property
    def m(self):
        """van Genuchten m = 1 - 1/n."""
        return 1.0 - (1.0 / self._n)

    def __str__(self):
        return (f" SoilProperties Id({id(self)}): \n n={self._n} \n alpha={self._alpha} \n"
                f" psi_sat={self._psi_sat} \n epsilon={self._epsilon}")
