"""Root density profile and water-uptake efficiency.

Mirrors the reference's ``TreeRoots`` (code/src/tree_roots.py:5-311): four root
density models normalised to unit integral, and the Lai-Katul / Collins-Bras
root efficiency used by the evapotranspiration sink.
"""
import math

import numpy as np

from .utilities import lerp_table


def _gamma_pdf(x, shape, scale):
    return x ** (shape - 1.0) * np.exp(-x / scale) / (scale ** shape * math.gamma(shape))


class TreeRoots:
    __slots__ = ("profile", "r_model", "max_depth_cm", "dz", "porous", "_knots")

    def __init__(self, ln, dz, r_model, porous):
        ln = int(ln)
        self.dz, self.r_model, self.porous = dz, r_model, porous
        self.max_depth_cm = ln * dz
        cells = np.linspace(1, 100, ln)          # depth mapped onto [1, 100]
        kind = str.upper(r_model)
        neg_exp = np.exp(-cells / 15) / 15
        if kind == "UNIFORM":
            pdf = np.ones(ln) / ln
        elif kind == "NEGATIVE_EXP":
            pdf = neg_exp
        elif kind == "GAMMA_PDF":
            pdf = _gamma_pdf(cells, 2.5, 5.0)
        elif kind == "MIXTURE":
            pdf = 0.15 * neg_exp + 0.85 * _gamma_pdf(cells, 2.5, 5.0)
        else:
            raise ValueError(f" {type(self).__name__}: Wrong root density profile type: {r_model}")
        pdf = np.maximum(pdf, 1.0e-8)
        self.profile = np.atleast_1d(pdf / (np.sum(pdf) * dz))
        self._knots = np.linspace(0.0, ln * dz, ln)

    @property
    def max_root_depth(self):
        return self.max_depth_cm

    @property
    def root_type(self):
        return self.r_model

    def __call__(self, z_new=None):
        if z_new is None:
            return self.profile
        return lerp_table(self._knots, self.profile, z_new)

    def efficiency(self, theta_z, z_roots):
        """Root efficiency rho(theta) [same shape as theta_z] and the water
        available for extraction in the root zone (tree_roots.py:179-292)."""
        theta_z, z_roots = np.atleast_1d(theta_z, z_roots)
        if theta_z.shape[0] != z_roots.shape[0]:
            raise RuntimeError(f" {type(self).__name__}: Input dimensions do not match: "
                               f"{theta_z.shape[0]} not equal to {z_roots.shape[0]}.")
        phi, f_cap, wlt = self.porous(z_roots)
        if theta_z.ndim == 2:
            phi, f_cap, wlt = (v[:, None] for v in (phi, f_cap, wlt))
        water = np.sum(theta_z - wlt, axis=0) * self.dz
        if not np.all(water > 0.0):
            return np.zeros(theta_z.shape), 0.0
        cum = np.cumsum(theta_z, axis=0) * self.dz
        total = np.atleast_1d(cum[-1]).copy()
        total[total == 0.0] = 1.0
        span = np.where(phi - wlt == 0.0, 1.0, phi - wlt)
        a1 = np.maximum(theta_z / span, cum / total)
        fspan = np.where(f_cap - wlt == 0.0, 1.0, f_cap - wlt)
        a2 = np.where((wlt < theta_z) & (theta_z <= f_cap), (theta_z - f_cap) / fspan, 0.0)
        a2 = np.where(theta_z > f_cap, 1.0, a2)
        a2 = np.clip(a2, 0.0, 1.0)
        if np.all(a2 == 1.0):
            a2 = a2 * 0.1
        rho = np.abs(a1 * a2)
        tot = np.atleast_1d(np.sum(rho, axis=0) * self.dz).copy()
        tot[tot == 0.0] = 1.0
        return rho / tot, water

    def __str__(self):
        return (f" Tree-Roots Id({id(self)}): \n Type={self.r_model} \n"
                f" Max-Depth={self.max_depth_cm} [cm].")
