"""
Deterministic synthetic inputs for the Richards time-stepping path.

The reference ships no data files (its CSVs are unpublished, README.md:10-13),
so every test / benchmark input is generated here.  The generators follow the
recipe fixed in SURVEY.md section 8(d) / BASELINE.md section 3:

* forcing rows ``ID, Datenum, Precipitation_cm, WTD_m`` in the column order the
  reference CLI reads (berkeley_hydro_main.py:115-116);
* per-column soil-parameter ensembles for configs 2-5 of BASELINE.json.

Nothing in here touches the GPU; it is plain NumPy host code.
"""
import copy

import numpy as np
import pandas as pd

FORCING_SEED = 20240607
ENSEMBLE_SEED = 1234
DATENUM_0 = 733682          # MATLAB datenum of 2008-10-01 (integer => hour 0)


def synthetic_forcing(n_rows=365, stiff=False, seed=FORCING_SEED,
                      half_hourly=False):
    """Return a DataFrame with the reference's four CSV columns.

    ``stiff=True`` is BASELINE.json config 5 (heavier, more frequent rain).
    ``half_hourly=True`` spaces the rows 1/48 day apart so that daylight
    windows (ET / surface evaporation) occur; the default is daily rows.
    """
    rng = np.random.default_rng(seed)
    d = np.arange(n_rows, dtype=np.float64)
    day = d / 48.0 if half_hourly else d
    wet = 0.5 * np.cos(2.0 * np.pi * (day - 105.0) / 365.0) + 0.5
    p_occ = 0.05 + 0.45 * wet + (0.2 if stiff else 0.0)
    u = rng.random(n_rows)
    amount = rng.gamma(0.7, 6.0 if stiff else 1.2, n_rows)
    if half_hourly:
        amount = amount / 8.0
    precip = np.where(u < p_occ, np.round(amount, 2), 0.0)
    wtd_m = -np.round(10.0 - 4.0 * wet, 4)
    return pd.DataFrame({"ID": np.arange(1, n_rows + 1),
                         "Datenum": DATENUM_0 + day,
                         "Precipitation_cm": precip,
                         "WTD_m": wtd_m})


def write_forcing_csv(frame, path):
    """Write the frame the way the reference CLI expects it (no header)."""
    frame.to_csv(path, header=False, index=False)


def default_params(site_information, **overrides):
    """The reference's default input_parameters.json as a dict
    (code/model_parameters/input_parameters.json:10-67), with the site file
    path made explicit.  ``overrides`` are merged one level deep."""
    params = {
        "IC_Filename": None,
        "Data_Filename": None,
        "Output_Name": "Sim_00",
        "Well_No": 10,
        "Site_Information": str(site_information),
        "Water_Content": {"Theta_Max": 0.30, "Theta_Min": 0.05,
                          "Theta_Residual": 0.04, "Wilting_Point_cm": -1500.0,
                          "Field_Capacity_cm": 340.0},
        "Hydraulic_Conductivity": {"Sat_Soil": 8.50, "Sat_Saprolite": 3.20,
                                   "Sat_Fresh_Bedrock": 0.10,
                                   "Sigma_Noise": 2.0, "Lambda_Exponent": 1.0},
        "Environmental": {"Interception_pct": 0.1, "Evaporation_pct": 0.0,
                          "Wet_Season_pct": 1.0, "Atmospheric_Demand": 0.5},
        "Trees": {"Leaf_Area_Index": 4.0, "Max_Root_Depth_cm": 1000.0,
                  "Root_Pdf_Profile": "Negative_Exp"},
        "Hydrological_Model": {"Name": "vrettas_fung",
                               "Porosity_Profile": "Stratified"},
        "Soil_Properties": {"n": 2.0, "a0": 0.009, "psi_sat": -0.0047,
                            "epsilon": 1.0e-7},
        "Simulation_Flags": {"SPINUP": False, "ET": True, "LF": True,
                             "HLIFT": False, "PREDICT": False},
    }
    for key, val in overrides.items():
        if isinstance(val, dict) and isinstance(params.get(key), dict):
            params[key].update(val)
        else:
            params[key] = val
    return params


def _log_uniform(rng, lo, hi):
    return float(np.exp(rng.uniform(np.log(lo), np.log(hi))))


def _member_rng(column, seed):
    """Generator of ensemble column ``column``: child ``column`` of
    ``SeedSequence(seed).spawn(...)``, built directly from its spawn key (the
    stream does not depend on how many columns the ensemble has)."""
    return np.random.default_rng(np.random.SeedSequence(seed, spawn_key=(int(column),)))


_MODELS = ("vrettas_fung", "vanGenuchten")
_PROFILES = ("Constant", "Linear", "Exponential", "Stratified")


def _member_draws(rng, mixed, perturb_theta):
    """Draw order per column stream: n, a0, psi_sat, [Theta_Max, precip
    multiplier], [model, profile]."""
    n = float(rng.uniform(1.6, 2.6))
    a0 = _log_uniform(rng, 0.005, 0.015)
    psi_sat = -_log_uniform(rng, 1.0e-3, 1.0)
    theta_max, mult, model, profile = None, 1.0, None, None
    if perturb_theta:
        theta_max = float(rng.uniform(0.25, 0.35))
        mult = float(np.exp(0.2 * rng.standard_normal()))
    if mixed:
        model = int(rng.integers(2))
        profile = int(rng.integers(4))
    return n, a0, psi_sat, theta_max, mult, model, profile


def ensemble_member(base_params, column, n_columns=None, mixed=False,
                    perturb_theta=False, seed=ENSEMBLE_SEED):
    """Parameters of ensemble column ``column``.

    Returns ``(params, precip_multiplier)``.  ``n_columns`` is accepted for
    compatibility and ignored: a member depends on its global id only.
    The wilting point is set to +1500 cm: the reference's field-capacity
    formula goes complex for non-even ``n`` with the default -1500
    (porosity.py:172-181).
    """
    n, a0, psi_sat, theta_max, mult, model, profile = \
        _member_draws(_member_rng(column, seed), mixed, perturb_theta)
    p = copy.deepcopy(base_params)
    p["Soil_Properties"]["n"] = n
    p["Soil_Properties"]["a0"] = a0
    p["Soil_Properties"]["psi_sat"] = psi_sat
    p["Soil_Properties"]["epsilon"] = 1.0e-7
    p["Water_Content"]["Wilting_Point_cm"] = 1500.0
    if perturb_theta:
        p["Water_Content"]["Theta_Max"] = theta_max
    if mixed:
        p["Hydrological_Model"]["Name"] = _MODELS[model]
        p["Hydrological_Model"]["Porosity_Profile"] = _PROFILES[profile]
    return p, mult


def ensemble_columns(base_params, first_column, count, mixed=False,
                     perturb_theta=False, seed=ENSEMBLE_SEED):
    """The same members as ``ensemble_member`` for the global ids
    ``first_column .. first_column + count - 1`` in struct-of-arrays form:
    ``(base, columns)`` where ``base`` is the parameter dict shared by all
    members (wilting point at +1500) and ``columns`` maps the per-column keys of
    ``pack.pack_ensemble`` to arrays of length ``count``."""
    base = copy.deepcopy(base_params)
    base["Soil_Properties"]["epsilon"] = 1.0e-7
    base["Water_Content"]["Wilting_Point_cm"] = 1500.0
    cols = {"n": np.empty(count), "a0": np.empty(count), "psi_sat": np.empty(count),
            "precip_mult": np.ones(count)}
    if perturb_theta:
        cols["Theta_Max"] = np.empty(count)
    if mixed:
        cols["model"] = np.empty(count, dtype=object)
        cols["profile"] = np.empty(count, dtype=object)
    for k in range(count):
        n, a0, psi_sat, theta_max, mult, model, profile = \
            _member_draws(_member_rng(first_column + k, seed), mixed, perturb_theta)
        cols["n"][k], cols["a0"][k], cols["psi_sat"][k] = n, a0, psi_sat
        cols["precip_mult"][k] = mult
        if perturb_theta:
            cols["Theta_Max"][k] = theta_max
        if mixed:
            cols["model"][k], cols["profile"][k] = _MODELS[model], _PROFILES[profile]
    return base, cols


def fine_grid_site(path, max_depth=5000.0):
    """Write a one-well site file whose column has >= 1001 nodes at the
    reference's hard-coded dz = 5 cm (config 5; simulation.py:101,140)."""
    import json
    site = {"Well": {"10": {"soil": 0.0, "saprolite": 50.0, "weathered": 200.0,
                            "max_depth": float(max_depth), "sat_depth": 125.0}}}
    with open(path, "w") as fh:
        json.dump(site, fh)
    return path
