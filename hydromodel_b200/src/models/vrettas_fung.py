"""Vrettas-Fung closure, host (NumPy) evaluation at arbitrary depths.

Same call signature and return tuple as the reference's ``VrettasFung``
(code/src/models/vrettas_fung.py:51-258): ``(theta, K, C, K_bkg, q_inf_max)``.
The solver does not call this: its closures are the device functions in
hydromodel_b200/csrc/hb_device.cuh, fed by the per-cell tables of pack.py.
"""
import numpy as np

from .hydrological_model import HydrologicalModel
from ..utilities import logN_rnd


class VrettasFung(HydrologicalModel):

    def __call__(self, psi, z, *args):
        from ...pack import layer_mean_k
        z, psi, phi = self._porosity_at(z, psi)
        n_rnd = np.atleast_1d(args[0]["n_rnd"])
        q, s_eff, sat, delta_s = self._retention(psi, phi)
        k = self.k_hc
        kmean, coef = layer_mean_k(z, self.porous.layers, k)
        # the noise index is local to the array handed in (reference :132-134,154)
        e = np.where(coef >= 0.0, coef, 0.0) * n_rnd[:z.size]
        if psi.ndim == 2:
            kmean, coef, e = (np.repeat(v[:, None], psi.shape[1], axis=1) for v in (kmean, coef, e))
        noisy = logN_rnd(kmean, k.sigma_noise * (1.0 - s_eff), e).reshape(s_eff.shape)
        k_bkg = np.where(coef >= 0.0, noisy, k.sat_soil)
        K = (s_eff ** k.lambda_exponent) * k_bkg
        K = np.minimum(np.where(sat, k_bkg, K), k_bkg)
        C = self._capacity(psi, s_eff, sat, delta_s)
        q_inf_max = np.minimum(2.0 * (phi[0] - q[0]) * self.dz, k_bkg[0])
        return q, K, C, k_bkg, q_inf_max
