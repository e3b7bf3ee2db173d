"""van Genuchten - Mualem closure, host (NumPy) evaluation.

Same call signature and return tuple as the reference's ``vanGenuchten``
(code/src/models/vanGenuchten.py:23-126); note the reference's Mualem term is
raised to the power ``n`` (not 2), which is reproduced.
"""
import numpy as np

from .hydrological_model import HydrologicalModel


class vanGenuchten(HydrologicalModel):

    def __call__(self, psi, z, *args):
        z, psi, phi = self._porosity_at(z, psi)
        q, s_eff, sat, delta_s = self._retention(psi, phi)
        k_sat = self.k_hc.sat_soil * np.ones(psi.shape)
        K = k_sat * np.sqrt(s_eff) * (1.0 - (1.0 - s_eff ** (1.0 / self.m)) ** self.m) ** self.n
        K = np.minimum(np.where(sat, k_sat, K), k_sat)
        C = self._capacity(psi, s_eff, sat, delta_s)
        q_inf_max = np.minimum(2.0 * (phi[0] - q[0]) * self.dz, k_sat[0])
        return q, K, C, k_sat, q_inf_max
