"""Shared helpers of the parity tests."""
import json
import os

import numpy as np

from hydromodel_b200 import pack
from hydromodel_b200.synthetic import synthetic_forcing, default_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SITE = os.path.join(ROOT, "hydromodel_b200", "model_parameters", "site_information.json")

_FORCING = None


def forcing365():
    global _FORCING
    if _FORCING is None:
        _FORCING = synthetic_forcing(365)
    return _FORCING


def column_from_json(params_json):
    params = json.loads(str(params_json))
    params["Site_Information"] = SITE
    return pack.pack_column(params, forcing365())


def default_column():
    return pack.pack_column(default_params(SITE), forcing365())


def tol_scaled(y, ref, rtol=1e-3, atol=1e-3):
    """Max error in units of the solver's own tolerance atol + rtol*|y|."""
    return float(np.max(np.abs(y - ref) / (atol + rtol * np.abs(ref))))


def rhs_cases(g, cfg):
    names = sorted({k.split("/")[1] for k in g.files if k.startswith(cfg + "/case")})
    for cs in names:
        p = f"{cfg}/{cs}/"
        precip, wtd, spin = g[p + "scal"]
        yield cs, g[p + "psi"], g[p + "n_rnd"], float(precip), int(wtd), int(spin), p


def rhs_configs(g):
    return sorted({k.split("/")[0] for k in g.files})
