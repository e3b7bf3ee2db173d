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
    stamps = pd.DatetimeIndex(stamps)
    rounded = stamps.round(freq="s")
    times = list(rounded)
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
    months = np.asarray(stamps.month)
    dry = np.isin(months, (4, 5, 6, 7, 8, 9))
    total = 13.4253 * 10 * env["Atmospheric_Demand"]
    et_pct = np.minimum(np.maximum(env["Wet_Season_pct"], 0.0), 1.0)
    n_dry = int(dry.sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        atm = np.where(dry, 2.0 * (1.0 - et_pct) * total / np.float64(n_dry),
                       2.0 * et_pct * total / np.float64(dim_t - n_dry))
    atm[np.isnan(atm)] = 0.0
    hours = np.asarray(rounded.hour)
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


# ---------------------------------------------------------------------------
# Ensembles: the same packing for many columns at once (array operations over
# the column axis; per-column Python work is O(1) scalars only).
# ---------------------------------------------------------------------------
ENSEMBLE_KEYS = {          # per-column key -> (parameter group, name in the JSON)
    "n": ("Soil_Properties", "n"), "a0": ("Soil_Properties", "a0"),
    "psi_sat": ("Soil_Properties", "psi_sat"), "epsilon": ("Soil_Properties", "epsilon"),
    "Theta_Max": ("Water_Content", "Theta_Max"), "Theta_Min": ("Water_Content", "Theta_Min"),
    "Theta_Residual": ("Water_Content", "Theta_Residual"),
    "model": ("Hydrological_Model", "Name"), "profile": ("Hydrological_Model", "Porosity_Profile"),
    "precip_mult": (None, None),
}
_PROFILE_KINDS = ("CONSTANT", "LINEAR", "EXPONENTIAL", "STRATIFIED")


def member_params(base_params, columns, k):
    """The ``input_parameters`` dict of column ``k`` of a struct-of-arrays
    ensemble and its precipitation multiplier."""
    import copy
    p = copy.deepcopy(base_params)
    for key, arr in columns.items():
        group, name = ENSEMBLE_KEYS[key]
        if group is not None:
            v = arr[k]
            p[group][name] = v if isinstance(v, str) else float(v)
    mult = float(columns["precip_mult"][k]) if "precip_mult" in columns else 1.0
    return p, mult


def _retention_scalar(alpha, n, m, suction):
    """(1 + (alpha*psi)**n)**m with the scalar arithmetic of the reference
    (porosity.py:172-181: Python/NumPy scalar pow); None when it goes complex."""
    try:
        d = (1.0 + (alpha * suction) ** n) ** m
    except (OverflowError, ZeroDivisionError):
        return None
    if isinstance(d, complex) or not np.isfinite(d):
        return None
    return float(d)


def pack_ensemble(base_params, data, columns, quiet=True, block=8192, _base=None):
    """Pack ``C`` columns that differ from ``base_params`` only in the keys of
    ``ENSEMBLE_KEYS`` (``columns``: key -> array of length C).  Returns the same
    arrays as ``C`` calls of ``pack_column`` would -- bit for bit, see
    tests/test_host_objects.py -- as ``par[C, NPARAM]``, ``stat[C, NSTATIC, nx]``
    plus the shared grid / forcing / objects of the base column.

    Columns whose parameter groups are invalid (the reference then falls back
    to defaults, simulation.py:100-231) or whose field-capacity formula leaves
    the reals (quirk Q7) are packed one by one through ``pack_column``.
    """
    unknown = set(columns) - set(ENSEMBLE_KEYS)
    if unknown:
        raise KeyError(f" pack_ensemble: unsupported per-column keys {sorted(unknown)}")
    sizes = {np.asarray(v).shape[0] for v in columns.values()}
    if len(sizes) != 1:
        raise ValueError(" pack_ensemble: per-column arrays must share one length.")
    C = sizes.pop()
    base = _base if _base is not None else pack_column(base_params, data, quiet=quiet)
    objs, forcing = base["objs"], base["forcing"]
    if C > block:                      # bounded temporaries: pack block by block
        nx_ = objs["z_grid"].size
        par, stat, n_slow = np.empty((C, NPARAM)), np.empty((C, NSTATIC, nx_)), 0
        for lo in range(0, C, block):
            part = pack_ensemble(base_params, data, {k: np.asarray(v)[lo:lo + block] for k, v in columns.items()},
                                 quiet=quiet, block=block, _base=base)
            par[lo:lo + block], stat[lo:lo + block] = part["par"], part["stat"]
            n_slow += part["n_slow"]
        return {"objs": objs, "forcing": forcing, "par": par, "stat": stat,
                "x": objs["z_grid"].astype(np.float64), "n_slow": n_slow}
    z, nx = objs["z_grid"], objs["z_grid"].size
    k_hc, theta0, soil0 = objs["K"], objs["theta"], objs["soil"]
    sp, wc = base_params["Soil_Properties"], base_params["Water_Content"]

    def col(key, default):
        if key in columns:
            return np.asarray(columns[key], dtype=np.float64)
        return np.full(C, float(default))

    n, a0 = col("n", sp["n"]), col("a0", sp["a0"])
    psi_sat, eps = col("psi_sat", sp["psi_sat"]), col("epsilon", sp["epsilon"])
    th_max = np.maximum(np.minimum(1.0, col("Theta_Max", wc["Theta_Max"])), 0.0)
    th_min = np.maximum(np.minimum(1.0, col("Theta_Min", wc["Theta_Min"])), 0.0)
    th_res = np.maximum(np.minimum(1.0, col("Theta_Residual", wc["Theta_Residual"])), 0.0)
    mult = col("precip_mult", 1.0)
    flc, wlt = wc["Field_Capacity_cm"], wc["Wilting_Point_cm"]
    hm = base_params["Hydrological_Model"]
    kinds = np.array([str.upper(str(v)) for v in columns["profile"]]) if "profile" in columns \
        else np.full(C, str.upper(hm["Porosity_Profile"]))
    model_id = np.array([0.0 if str.upper(str(v)) == "VRETTAS_FUNG" else 1.0 for v in columns["model"]]) \
        if "model" in columns else np.full(C, float(objs["model_id"]))

    # columns the array path cannot reproduce: invalid groups, NOISY, complex field capacity
    slow = ~((n > 1.0) & (a0 > 0.0) & (0.0 <= th_res) & (th_res < th_min) & (th_min < th_max) &
             (th_max <= 1.0) & np.isin(kinds, _PROFILE_KINDS))
    m = 1.0 - (1.0 / n)
    d_flc, d_wlt = np.ones(C), np.ones(C)
    for k in range(C):
        if slow[k]:
            continue
        a, b = _retention_scalar(float(a0[k]), float(n[k]), float(m[k]), flc), \
            _retention_scalar(float(a0[k]), float(n[k]), float(m[k]), wlt)
        if a is None or b is None:
            slow[k] = True
        else:
            d_flc[k], d_wlt[k] = a, b

    # ---- porosity profiles at the nodes, by profile type (porosity.py:22-185) ----
    l0, l_sapr, l_wbed, l_fbed = objs["layers"]
    in_sapr = (z >= l_sapr) & (z <= l_wbed)
    in_wbed = (z >= l_wbed) & (z <= l_fbed)
    th_mid = 0.5 * (th_max + th_min)
    q_sat = np.empty((C, nx))
    for kind in _PROFILE_KINDS:
        sel = np.flatnonzero((kinds == kind) & ~slow)
        if not sel.size:
            continue
        mx, mn, md = th_max[sel, None], th_min[sel, None], th_mid[sel, None]
        if kind == "CONSTANT":
            q = mx * np.ones(nx)
        elif kind == "LINEAR":
            q = np.linspace(th_max[sel], th_min[sel], nx, axis=1)
        elif kind == "EXPONENTIAL":
            q = mx * np.exp(-z[None, :] * (np.log(mx / mn) / z[-1]))
        else:
            q = mx * np.ones(nx)
            if np.any(in_sapr):
                q[:, in_sapr] = np.linspace(th_max[sel], th_mid[sel], int(in_sapr.sum()), axis=1)
            if np.any(in_wbed):
                rate = np.log(md / mn) / z[-1]
                q[:, in_wbed] = md * np.exp(-np.linspace(0, l_fbed, int(in_wbed.sum()))[None, :] * rate)
        q_sat[sel] = np.minimum(np.maximum(q, mn), mx)
    res = th_res[:, None]
    with np.errstate(all="ignore"):
        f_cap = np.maximum(res + (q_sat - res) / d_flc[:, None], res)
        wilt = np.minimum(res + (q_sat - res) / d_wlt[:, None], f_cap)

    # ---- static cell arrays: shared rows from the base column, porosity rows per column ----
    stat = np.empty((C, NSTATIC, nx))
    stat[:] = base["stat"][None]
    x_mid = z[:-1] + 0.5 * np.diff(z)
    prof_mid = lerp_table(z, np.stack((q_sat, f_cap, wilt), axis=1), x_mid)     # [C, 3, nx-1]
    stat[:, S_PHI_MID, :-1] = prof_mid[:, 0]
    stat[:, S_FCAP_MID, :-1] = prof_mid[:, 1]
    stat[:, S_WILT_MID, :-1] = prof_mid[:, 2]
    stat[:, S_PHI_MID, -1] = q_sat[:, -1]
    stat[:, S_PHI_NODE] = q_sat

    # ---- scalar blocks ----
    par = np.empty((C, NPARAM))
    par[:] = base["par"][None]
    par[:, P_N], par[:, P_M], par[:, P_ALPHA] = n, m, a0
    par[:, P_PSI_SAT] = np.minimum(psi_sat, 0.0)
    par[:, P_EPS] = np.maximum(np.maximum(eps, 1.0e-8), 1.0e-8)
    par[:, P_THETA_RES] = th_res
    par[:, P_MODEL], par[:, P_PRECIP_MULT] = model_id, mult
    par[:, P_PHI_TOP] = q_sat[:, 0]
    ds = q_sat - res
    if z[0] == 0.0:
        par[:, P_DSMIN_ALL] = ds.min(axis=1)
        par[:, P_DSMIN_EXCL0] = ds[:, 1:].min(axis=1) if nx > 1 else ds[:, 0]
    else:
        par[:, P_DSMIN_ALL] = par[:, P_DSMIN_EXCL0] = ds[:, 0]
    theta_50 = np.maximum(0.5 ** (1.0 / k_hc.lambda_exponent), 0.05) \
        if k_hc.lambda_exponent > 0 else 0.05
    par[:, P_IPSI50] = a0 / np.sqrt(theta_50 ** (-2.0) - 1.0)

    for k in np.flatnonzero(slow):
        p, mk = member_params(base_params, columns, k)
        one = pack_column(p, data, precip_mult=mk, quiet=quiet)
        par[k], stat[k] = one["par"], one["stat"]
    return {"objs": objs, "forcing": forcing, "par": par, "stat": stat,
            "x": z.astype(np.float64), "n_slow": int(slow.sum())}


def split_params_list(params_list):
    """Group a list of ``input_parameters`` dicts by everything that is NOT a
    per-column key of ``pack_ensemble``: returns ``[(base, columns, indices)]``
    so that each group can be packed with array operations."""
    groups = {}
    for idx, p in enumerate(params_list):
        shared = {g: ({k: v for k, v in blk.items()
                       if (g, k) not in _ENSEMBLE_PAIRS} if isinstance(blk, dict) else blk)
                  for g, blk in p.items()}
        sig = json.dumps(shared, sort_keys=True, default=str)
        groups.setdefault(sig, []).append(idx)
    out = []
    for idxs in groups.values():
        base = params_list[idxs[0]]
        cols = {}
        for key, (g, name) in ENSEMBLE_KEYS.items():
            if g is None:
                continue
            vals = [params_list[i][g][name] for i in idxs]
            if isinstance(vals[0], str):
                cols[key] = np.array(vals, dtype=object)
            else:
                cols[key] = np.array(vals, dtype=np.float64)
        out.append((base, cols, np.array(idxs)))
    return out


_ENSEMBLE_PAIRS = {v for v in ENSEMBLE_KEYS.values() if v[0] is not None}
