"""Porosity, field-capacity and wilting-point profiles of one column.

Mirrors the reference's ``Porosity`` (code/src/porosity.py:7-230): five profile
types built on the node grid, evaluated elsewhere by linear interpolation.  In
this package the profiles are static inputs of the device solver: they are
sampled once per column at the nodes and the cell midpoints
(hydromodel_b200/pack.py) instead of being interpolated on every right-hand
side call.
"""
import numpy as np

from .utilities import logN_rnd, lerp_table


class Porosity:
    __slots__ = ("profile", "field_cap", "wilting_point", "p_model",
                 "p_layers", "_z")

    def __init__(self, z_grid, layers, theta, soil, p_model):
        z = np.atleast_1d(z_grid)
        if not z.size:
            raise ValueError(f" {type(self).__name__}: Input array z_grid is empty.")
        if np.any(np.diff(z) <= 0.0):
            raise RuntimeError(f" {type(self).__name__}: Space domain z_grid is not increasing.")
        self.p_model, self.p_layers, self._z = p_model, layers, z
        _, l_sapr, l_wbed, l_fbed = layers
        in_sapr = (z >= l_sapr) & (z <= l_wbed)
        in_wbed = (z >= l_wbed) & (z <= l_fbed)
        kind = str.upper(p_model)

        def bedrock_decay(count):
            # exponential from theta.mid towards theta.min over the bedrock
            rate = np.log(theta.mid / theta.min) / z[-1]
            return theta.mid * np.exp(-np.linspace(0, l_fbed, count) * rate)

        if kind == "CONSTANT":
            q_sat = theta.max * np.ones(z.size)
        elif kind == "LINEAR":
            q_sat = np.linspace(theta.max, theta.min, z.size)
        elif kind == "EXPONENTIAL":
            q_sat = theta.max * np.exp(-z * (np.log(theta.max / theta.min) / z[-1]))
        elif kind in ("STRATIFIED", "NOISY"):
            q_sat = theta.max * np.ones(z.size)
            rng = np.random.default_rng() if kind == "NOISY" else None
            if np.any(in_sapr):
                mean = np.linspace(theta.max, theta.mid, in_sapr.sum())
                q_sat[in_sapr] = mean if rng is None else logN_rnd(
                    mean, 5.0e-5 * np.min(mean), rng.standard_normal(mean.size))
            if np.any(in_wbed):
                mean = bedrock_decay(in_wbed.sum())
                q_sat[in_wbed] = mean if rng is None else logN_rnd(
                    mean, 5.0e-5 * np.min(mean), rng.standard_normal(mean.size))
        else:
            raise ValueError(f" {type(self).__name__}: Wrong porosity profile type: {p_model}")

        q_sat = np.minimum(np.maximum(q_sat, theta.min), theta.max)
        self.profile = q_sat.flatten()

        def retention(psi):
            return theta.res + (q_sat - theta.res) / (1.0 + (soil.alpha * psi) ** soil.n) ** soil.m

        self.field_cap = np.maximum(retention(theta.flc), theta.res)
        self.wilting_point = np.minimum(retention(theta.wlt), self.field_cap)

    def __call__(self, z_new=None):
        """(porosity, field capacity, wilting point) at ``z_new``; the full
        node profiles when ``z_new`` is None or all-zero (the reference tests
        ``np.any(z_new)``, porosity.py:200)."""
        if np.any(z_new):
            return lerp_table(self._z, np.stack((self.profile, self.field_cap,
                                                 self.wilting_point)), z_new)
        return self.profile, self.field_cap, self.wilting_point

    @property
    def layers(self):
        return self.p_layers

    def __str__(self):
        return f" Porosity Id({id(self)}): Type={self.p_model}"
