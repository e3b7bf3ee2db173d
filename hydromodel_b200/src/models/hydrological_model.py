"""Base class of the hydrological closures (host side).

Mirrors the reference's ``HydrologicalModel``
(code/src/models/hydrological_model.py:4-121): captures the soil parameters
(``epsilon`` floored at 1e-8) and provides the inverse retention curve
``pressure_head(theta, z)`` that seeds the spin-up.  The forward closures used
by the time integrator run on the device (hydromodel_b200/csrc/hb_device.cuh);
the NumPy ``__call__`` of the subclasses is the same map for arbitrary depths,
kept for API compatibility of direct calls.
"""
import numpy as np


class HydrologicalModel:
    def __init__(self, soil, porous, k_hc, theta_res, dz):
        self.n, self.m, self.alpha = soil.n, soil.m, soil.alpha
        self.psi_sat = soil.psi_sat
        self.epsilon = np.maximum(soil.epsilon, 1.0e-8)
        self.k_hc, self.porous = k_hc, porous
        self.dz, self.theta_res = dz, theta_res

    # -- helpers shared by the subclasses ---------------------------------
    def _porosity_at(self, z, psi):
        """Porosity at z, shaped like psi (column-replicated for 2-D psi)."""
        z, psi = np.atleast_1d(z, psi)
        if psi.shape[0] != z.shape[0]:
            raise ValueError(f" {type(self).__name__}: Input size dimensions don't match:"
                             f" {psi.shape[0]} not equal to {z.shape[0]}.")
        phi = np.atleast_1d(self.porous(z)[0])
        if psi.ndim == 2:
            phi = np.repeat(phi[:, None], psi.shape[1], axis=1)
        return z, psi, phi

    def _retention(self, psi, phi):
        """theta, effective saturation, the saturated mask and phi - theta_res."""
        delta_s = phi - self.theta_res
        sat = psi >= self.psi_sat
        q = self.theta_res + delta_s * (1.0 + (self.alpha * np.abs(psi)) ** self.n) ** (-self.m)
        q = np.where(sat, phi, q)
        s_eff = np.clip((q - self.theta_res) / delta_s, 0.0, 1.0)
        return q, s_eff, sat, delta_s

    def _capacity(self, psi, s_eff, sat, delta_s):
        with np.errstate(all="ignore"):
            c = (self.m * self.n) * self.alpha * delta_s * \
                (s_eff ** (1.0 / self.m + 1.0)) * (self.alpha * np.abs(psi)) ** (self.n - 1.0)
        c = np.where(sat, self.epsilon, c)
        return np.where((c < self.epsilon) | ~np.isfinite(c), self.epsilon, c)

    def pressure_head(self, theta, z):
        """psi(theta) at depths z and the effective saturation."""
        z, theta = np.atleast_1d(z, theta)
        if theta.shape[0] != z.shape[0]:
            raise ValueError(f" {type(self).__name__}: Input size dimensions don't match:"
                             f" {theta.shape[0]} not equal to {z.shape[0]}.")
        phi = np.atleast_1d(self.porous(z)[0])
        if theta.ndim == 2:
            phi = np.repeat(phi[:, None], theta.shape[1], axis=1)
        q = np.minimum(np.maximum(theta, self.theta_res), phi)
        s_eff = np.minimum(np.maximum((q - self.theta_res) / (phi - self.theta_res),
                                      self.epsilon), 1.0)
        wet = s_eff >= 0.99998
        psi = np.zeros(theta.shape)
        psi[~wet] = -((s_eff[~wet] ** (-1.0 / self.m) - 1.0) ** (1.0 / self.n)) / self.alpha
        psi[wet] = np.arange(0, np.sum(wet)) * self.dz
        psi[~np.isfinite(psi)] = -1.0e+5
        return psi, s_eff
