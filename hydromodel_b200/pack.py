"""
Host precompute layer: turn the reference's inputs (one ``input_parameters``
dict + one forcing table) into the flat FP64 arrays the device solver consumes
(layout: include/hb_layout.h).

Everything the reference re-derives on every right-hand-side call from static
objects is evaluated here once per column:

* porosity at nodes and midpoints      (reference porosity.py:184-206)
* layer-mean background conductivity   (reference vrettas_fung.py:22-49,132-190)
* layer noise coefficients             (reference vrettas_fung.py:154,172,190)
* grid-snapped water-table observations(reference simulation.py:245-255,579)
* daylight flags / transpiration series(reference simulation.py:239-242,273-333)
* finite-difference colouring          (scipy.optimize group_columns as called
                                        by scipy BDF, bdf.py:263-268)
"""
import json
from pathlib import Path

import numpy as np
import pandas as pd

from .src.hydraulic_conductivity import HydraulicConductivity
from .src.porosity import Porosity
from .src.soil_properties import SoilProperties
from .src.tree_roots import TreeRoots
from .src.utilities import lerp_table
from .src.water_content import WaterContent

# ---- include/hb_layout.h mirrored -----------------------------------------
(P_N, P_M, P_ALPHA, P_PSI_SAT, P_EPS, P_THETA_RES, P_LAMBDA, P_SIGMA,
 P_KSAT_SOIL, P_DZ, P_INTERCEPT, P_SURF_EVAP, P_MODEL, P_PRECIP_MULT,
 P_PHI_TOP, P_KMEAN_TOP, P_COEF_TOP, P_DSMIN_ALL, P_DSMIN_EXCL0, P_FLAGS,
 P_ROOT_TOP, P_LAI, P_IPSI50, P_RESERVED0, NPARAM) = range(25)
(S_PHI_MID, S_KMEAN_MID, S_COEF_MID, S_PHI_NODE, S_KMEAN_NODE, S_COEF_NODE,
 S_ROOT_MID, S_FCAP_MID, S_WILT_MID, NSTATIC) = range(10)
(W_WTD_EST, W_LATERAL, W_TRANSP, W_STATUS, NWSCAL) = range(5)
(C_NFEV, C_NJEV, C_NLU, C_NSTEPS, C_NREJECT, C_NFAIL, C_NRHS, C_NNEWTON,
 C_NWINDOWS, NCOUNTER) = range(10)
(F_PSI, F_THETA, F_K, F_KBKG, NFIELD) = range(5)
FLAG_LF, FLAG_ET, FLAG_HLIFT, FLAG_PREDICT = 1, 2, 4, 8
NOISE_TAPE, NOISE_PHILOX = 0, 1

DZ_CM = 5.0          # hard-coded in the reference (simulation.py:101)


def fd_groups(n):
    """Column colouring scipy's BDF uses for the tridiagonal sparsity pattern:
    greedy grouping over the ``RandomState(0)`` permutation
    (scipy/optimize/_numdiff.py:227-284, _group_columns.group_sparse).
    Returns ``(groups[n] int32, n_groups)``."""
    order = np.random.RandomState(0).permutation(n)
    rows = [tuple(r for r in (c - 1, c, c + 1) if 0 <= r < n) for c in order]
    grp = -np.ones(n, dtype=np.int64)
    current = 0
    for i in range(n):
        if grp[i] >= 0:
            continue
        grp[i] = current
        union = np.zeros(n, dtype=bool)
        union[list(rows[i])] = True
        all_grouped = True
        for j in range(n):
            if grp[j] >= 0:
                continue
            all_grouped = False
            if not any(union[r] for r in rows[j]):
                union[list(rows[j])] = True
                grp[j] = current
        if all_grouped:
            break
        current += 1
    out = np.empty(n, dtype=np.int32)
    out[order] = grp
    return out, int(out.max()) + 1


def load_site(params):
    with open(Path(params["Site_Information"]), "r") as fh:
        site = json.load(fh)
    key = str(params["Well_No"])
    if key not in site["Well"]:
        raise ValueError(" Simulation: The selected well does not exist in the"
                         " site information file.")
    well = site["Well"][key]
    if well["sat_depth"] >= well["max_depth"]:
        raise RuntimeError(" Simulation: The well seems fully saturated.")
    return well


def build_objects(params, quiet=True):
    """Construct the host objects the way Simulation.setupModel does
    (reference simulation.py:100-231), falling back to defaults when a
    parameter group is invalid."""
    well = load_site(params)
    dz = DZ_CM
    layers = (well["soil"], well["saprolite"], well["weathered"], well["max_depth"])
    z_grid = np.arange(well["soil"], well["max_depth"] + dz, dz)

    def guarded(cls, label, *a):
        try:
            return cls(*a)
        except Exception as err:          # noqa: BLE001 - reference semantics
            if not quiet:
                print(f" {label} failed to initialize: {err}."
                      f" It will use default initialization parameters.")
            return cls()

    sp, wc, hc = (params["Soil_Properties"], params["Water_Content"],
                  params["Hydraulic_Conductivity"])
    soil = guarded(SoilProperties, "SoilProperties",
                   sp["n"], sp["a0"], sp["psi_sat"], sp["epsilon"])
    theta = guarded(WaterContent, "WaterContent", wc["Theta_Min"], wc["Theta_Max"],
                    wc["Theta_Residual"], wc["Wilting_Point_cm"], wc["Field_Capacity_cm"])
    k_hc = guarded(HydraulicConductivity, "HydraulicConductivity", hc["Sat_Soil"],
                   hc["Sat_Saprolite"], hc["Sat_Fresh_Bedrock"], hc["Sigma_Noise"],
                   hc["Lambda_Exponent"])
    porous = Porosity(z_grid, layers, theta, soil,
                      params["Hydrological_Model"]["Porosity_Profile"])
    roots = TreeRoots(np.ceil(params["Trees"]["Max_Root_Depth_cm"] / dz), dz,
                      params["Trees"]["Root_Pdf_Profile"], porous)
    is_vf = str.upper(params["Hydrological_Model"]["Name"]) == "VRETTAS_FUNG"
    return {"well": well, "dz": dz, "layers": layers, "z_grid": z_grid,
            "sat_cells": np.ceil(well["sat_depth"] / dz), "soil": soil,
            "theta": theta, "K": k_hc, "porosity": porous, "tree_roots": roots,
            "model_id": 0 if is_vf else 1,
            "sim_flags": params["Simulation_Flags"],
            "env_param": params["Environmental"],
            "LAI": params["Trees"]["Leaf_Area_Index"]}


def layer_mean_k(z, layers, k_hc):
    """Layer-mean background conductivity and noise coefficient at depths z
    (reference vrettas_fung.py:26-48 for the 1-cm knot tables, :132-190 for
    the layer membership and the 0.05 / 0.10 / 1.0 noise scalings).  Cells
    outside every layer get coef = -1 (K_bkg stays Ksat_soil, :143)."""
    z = np.atleast_1d(np.asarray(z, dtype=float))
    l0, l1, l2, l3 = layers
    kmean = np.full(z.shape, float(k_hc.sat_soil))
    coef = np.full(z.shape, -1.0)
    soil = (z >= l0) & (z < l1)
    sapr = (z >= l1) & (z < l2)
    wbed = (z >= l2) & (z <= l3)
    coef[soil] = 0.05
    if np.any(sapr):
        knots = np.arange(l1, l2 + 1)
        vals = np.linspace(k_hc.sat_soil, k_hc.sat_saprolite, knots.size)
        kmean[sapr] = lerp_table(knots, vals, z[sapr])
        coef[sapr] = 0.10
    if np.any(wbed):
        knots = np.arange(l2, l3 + 1)
        p0 = k_hc.sat_saprolite
        p1 = np.log(p0 / k_hc.sat_fresh_bedrock) / l3
        vals = p0 * np.exp(-np.linspace(0, l3, knots.size) * p1)
        kmean[wbed] = lerp_table(knots, vals, z[wbed])
        coef[wbed] = 1.0
    return kmean, coef


def pack_static(objs, interception=0.0, surface_evap=0.0, precip_mult=1.0):
    """Per-column scalar block ``par[NPARAM]`` and static cell arrays
    ``stat[NSTATIC, nx]``."""
    z = objs["z_grid"]
    nx = z.size
    soil, theta, k_hc, porous = objs["soil"], objs["theta"], objs["K"], objs["porosity"]
    x_mid = z[:-1] + 0.5 * np.diff(z)
    stat = np.zeros((NSTATIC, nx))
    prof_mid = lerp_table(z, np.stack((porous.profile, porous.field_cap,
                                       porous.wilting_point)), x_mid)
    stat[S_PHI_MID, :-1] = prof_mid[0]
    stat[S_FCAP_MID, :-1] = np.real(prof_mid[1])
    stat[S_WILT_MID, :-1] = np.real(prof_mid[2])
    stat[S_PHI_MID, -1] = porous.profile[-1]
    stat[S_PHI_NODE] = porous.profile
    km, cf = layer_mean_k(x_mid, objs["layers"], k_hc)
    stat[S_KMEAN_MID, :-1], stat[S_COEF_MID, :-1] = km, cf
    stat[S_KMEAN_MID, -1], stat[S_COEF_MID, -1] = k_hc.sat_soil, -1.0
    stat[S_KMEAN_NODE], stat[S_COEF_NODE] = layer_mean_k(z, objs["layers"], k_hc)
    roots = objs["tree_roots"]
    in_root = x_mid <= roots.max_root_depth
    stat[S_ROOT_MID, :-1][in_root] = roots(x_mid[in_root])

    flags = objs["sim_flags"]
    bits = (FLAG_LF * bool(flags.get("LF")) | FLAG_ET * bool(flags.get("ET")) |
            FLAG_HLIFT * bool(flags.get("HLIFT")) | FLAG_PREDICT * bool(flags.get("PREDICT")))
    km_top, cf_top = layer_mean_k(z[:1], objs["layers"], k_hc)
    par = np.zeros(NPARAM)
    par[P_N], par[P_M], par[P_ALPHA] = soil.n, soil.m, soil.alpha
    par[P_PSI_SAT] = soil.psi_sat
    par[P_EPS] = np.maximum(soil.epsilon, 1.0e-8)
    par[P_THETA_RES] = theta.res
    par[P_LAMBDA], par[P_SIGMA] = k_hc.lambda_exponent, k_hc.sigma_noise
    par[P_KSAT_SOIL], par[P_DZ] = k_hc.sat_soil, objs["dz"]
    par[P_INTERCEPT], par[P_SURF_EVAP] = interception, surface_evap
    par[P_MODEL], par[P_PRECIP_MULT] = objs["model_id"], precip_mult
    par[P_PHI_TOP] = porous.profile[0]
    par[P_KMEAN_TOP], par[P_COEF_TOP] = km_top[0], cf_top[0]
    ds = porous.profile - theta.res
    # bc_fun evaluates the closure at z = x[0]; when that is 0.0 the porosity
    # call returns the whole profile and the evaporation test is taken over
    # the broadcast vector (reference porosity.py:200, richards_pde.py:461).
    if z[0] == 0.0:
        par[P_DSMIN_ALL] = ds.min()
        par[P_DSMIN_EXCL0] = ds[1:].min() if nx > 1 else ds[0]
    else:
        par[P_DSMIN_ALL] = par[P_DSMIN_EXCL0] = ds[0]
    par[P_FLAGS] = float(bits)
    par[P_ROOT_TOP] = float(x_mid[0] <= roots.max_root_depth)
    par[P_LAI] = objs["LAI"]
    theta_50 = np.maximum(0.5 ** (1.0 / k_hc.lambda_exponent), 0.05) \
        if k_hc.lambda_exponent > 0 else 0.05
    par[P_IPSI50] = soil.alpha / np.sqrt(theta_50 ** (-2.0) - 1.0)
    return par, stat


def pack_forcing(objs, params, data):
    """Per-row forcing arrays from the four-column table the reference CLI
    reads.  Returns a dict of arrays of length T (reference
    simulation.py:233-352)."""
    z = objs["z_grid"]
    stamps = pd.to_datetime(data.loc[:, "Datenum"] - 719529, unit="D")
    times = [t.round(freq="s") for t in stamps]
    z_wtd = np.array(np.abs(np.round(100.0 * data.loc[:, "WTD_m"])))
    if np.any(np.isnan(z_wtd)):
        raise RuntimeError(" Simulation: Water table depth observations contain NaN values.")
    if np.any(z_wtd > z[-1]):
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    snapped = z[np.searchsorted(z, z_wtd, side="left")]
    precip = np.array(data.loc[:, "Precipitation_cm"], dtype=float)
    if np.any(np.isnan(precip)):
        raise ValueError(" Simulation: Precipitation observations contain NaN values.")
    env = params["Environmental"]
    dim_t = len(times)
    months = np.array([t.month for t in stamps])
    dry = np.isin(months, (4, 5, 6, 7, 8, 9))
    total = 13.4253 * 10 * env["Atmospheric_Demand"]
    et_pct = np.minimum(np.maximum(env["Wet_Season_pct"], 0.0), 1.0)
    n_dry = int(dry.sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        atm = np.where(dry, 2.0 * (1.0 - et_pct) * total / np.float64(n_dry),
                       2.0 * et_pct * total / np.float64(dim_t - n_dry))
    atm[np.isnan(atm)] = 0.0
    hours = np.array([t.hour for t in times])
    # node index of each snapped observation (always on the grid after
    # snapping; kept general: -1 marks "skip window", simulation.py:579-588)
    idx = np.searchsorted(z, snapped)
    wtd_idx = np.where(z[np.minimum(idx, z.size - 1)] == snapped, idx, -1).astype(np.int32)
    return {"time": times, "zWtd_cm": snapped, "precipitation_cm": precip,
            "wtd_idx": wtd_idx,
            "daylight": ((hours >= 6) & (hours < 18)).astype(np.int32),
            "month": months.astype(np.int32), "atm": atm,
            "interception": float(env["Interception_pct"]),
            "surface_evap": float(2.0 * np.sum(env["Evaporation_pct"] * precip) / dim_t),
            "dim_t": dim_t}


def pack_column(params, data, precip_mult=1.0, quiet=True):
    """One JSON + one forcing table -> everything the device needs."""
    objs = build_objects(params, quiet=quiet)
    forcing = pack_forcing(objs, params, data)
    par, stat = pack_static(objs, forcing["interception"], forcing["surface_evap"],
                            precip_mult)
    return {"objs": objs, "forcing": forcing, "par": par, "stat": stat,
            "x": objs["z_grid"].astype(np.float64)}


def saturated_initial_state(objs):
    """The spin-up's starting pressure head: the porosity profile pushed through
    the inverse retention curve, i.e. a fully saturated column psi_k = k*dz
    (reference simulation.py:414-417, hydrological_model.py:91-115)."""
    from .src.models.hydrological_model import HydrologicalModel
    hm = HydrologicalModel(objs["soil"], objs["porosity"], objs["K"],
                           objs["theta"].res, objs["dz"])
    psi, _ = hm.pressure_head(objs["porosity"].profile, objs["z_grid"])
    return psi
