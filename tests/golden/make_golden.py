#!/usr/bin/env python3
"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (vrettasm/HydroModel mounted read-only at /root/reference) in this
container.  The reference cannot travel to the GPU box, so its outputs are
committed as small .npz files together with this script.

Stack the fixtures were produced on ("reference @ container stack"):
python 3.12.3, numpy 2.3.5, scipy 1.18.1, numba 0.65.0, pandas 3.0.2 (the
reference pins scipy 1.5.2 which is not available offline).

What is recorded
----------------
rhs_cases.npz      (psi, args) -> dpsi/dt of RichardsPDE.__call__
                   (richards_pde.py:82-160) plus the closure outputs on the
                   nodes (vrettas_fung.py:51-258 / vanGenuchten.py:23-126) for
                   a matrix of models / porosity profiles / soil parameters /
                   flags / states.
jac_cases.npz      scipy num_jac (common.py:268-455) tridiagonal Jacobian and
                   step factors at window start for two states.
windows_*.npz      teacher-forcing tapes of Simulation.run windows
                   (simulation.py:576-649): y0, args, y_end, solver counters
                   (nfev, njev, nlu, accepted steps, status, accepted times).
spinup_default.npz tape of Simulation.initial_conditions (simulation.py:389-493).

Usage:  python tests/golden/make_golden.py [--out tests/golden]
"""
import argparse
import os
import sys
import types
import warnings

import numpy as np

REF_CODE = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

sys.dont_write_bytecode = True
sys.modules.setdefault("h5py", types.ModuleType("h5py"))   # only saveResults needs it
sys.path.insert(0, REF_CODE)
warnings.filterwarnings("ignore", category=DeprecationWarning)

import pandas as pd                                            # noqa: E402
from scipy.integrate._ivp.common import num_jac                # noqa: E402
from scipy.optimize._numdiff import group_columns              # noqa: E402
from scipy import sparse                                       # noqa: E402

import src.richards_pde as ref_pde                             # noqa: E402
from src.simulation import Simulation                          # noqa: E402

from hydromodel_b200.synthetic import (synthetic_forcing, default_params,
                                       ensemble_member, fine_grid_site)   # noqa: E402

SITE = os.path.join(REF_CODE, "model_parameters", "site_information.json")

# ---------------------------------------------------------------------------
# Instrumentation: wrap scipy's solve_ivp as seen by the reference module so
# that each call's counters are taped.  The reference source is untouched.
# ---------------------------------------------------------------------------
_TAPE = []
_real_solve_ivp = ref_pde.solve_ivp


def _taping_solve_ivp(fun, t_span, y0, **kw):
    args = kw["args"][0]
    rec = {"y0": np.array(y0, dtype=float), "n_rnd": np.array(args["n_rnd"]),
           "wtd": int(args["wtd"]), "precip": float(args["precipitation"]),
           "t0": float(t_span[0]), "atm": float(args.get("atm", 0.0)),
           "hour": int(args["time"].hour) if "time" in args else 0}
    sol = _real_solve_ivp(fun, t_span=t_span, y0=y0, **kw)
    rec.update(y_end=sol.y[:, -1].copy(), nfev=sol.nfev, njev=sol.njev,
               nlu=sol.nlu, nsteps=sol.t.size - 1, status=sol.status,
               t_acc=sol.t.copy(),
               lateral_flow=float(fun.arg_out["lateral_flow"]),
               transpiration=float(fun.arg_out["transpiration"]))
    _TAPE.append(rec)
    return sol


ref_pde.solve_ivp = _taping_solve_ivp


def make_sim(params, forcing, seed, ic=None):
    """Build a reference Simulation; with ``ic`` the spin-up is bypassed by
    handing the reference an IC file exactly as its IC_Filename path does
    (simulation.py:358-385)."""
    sim = Simulation("golden", seed=seed)
    if ic is not None:
        import tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as fh:
            for v in ic:
                fh.write(repr(float(v)) + "\n")
        params = dict(params, IC_Filename=fh.name)
    sim.setupModel(params, forcing)
    return sim


def static_pack(sim):
    """Everything a restatement needs to rebuild the column."""
    m = sim.mData
    por = m["porosity"]
    return {"z": m["z_grid"], "porosity": por.profile,
            "field_cap": por.field_cap, "wilting": por.wilting_point,
            "layers": np.array(m["layers"], dtype=float),
            "zwtd_cm": m["zWtd_cm"], "precip_cm": m["precipitation_cm"],
            "atm_series": m["atm"], "surface_evap": float(m["surface_evap"]),
            "interception": float(m["interception"])}


def tape_to_arrays(tape, prefix=""):
    out = {}
    for key in ("y0", "n_rnd", "y_end"):
        out[prefix + key] = np.stack([r[key] for r in tape])
    for key in ("wtd", "nfev", "njev", "nlu", "nsteps", "status"):
        out[prefix + key] = np.array([r[key] for r in tape], dtype=np.int64)
    for key in ("precip", "t0", "lateral_flow", "transpiration", "atm", "hour"):
        out[prefix + key] = np.array([r[key] for r in tape], dtype=float)
    out[prefix + "t_acc"] = np.concatenate([r["t_acc"] for r in tape])
    out[prefix + "t_acc_off"] = np.cumsum([0] + [r["t_acc"].size for r in tape])
    return out


# ---------------------------------------------------------------------------
def gen_default_run(out_dir, n_windows):
    """Default JSON, seed 1234: spin-up tape + first n_windows of run()."""
    forcing = synthetic_forcing(365)
    params = default_params(SITE)
    _TAPE.clear()
    sim = make_sim(params, forcing, seed=1234)          # runs the spin-up
    spin = list(_TAPE)
    print(f"  spin-up: {len(spin)} solves")
    # keep the first 24 and the last 8 spin-up solves (teacher forcing needs
    # y0 per solve, so any subset is self-contained), plus the final IC.
    keep = spin[:24] + spin[-8:]
    np.savez_compressed(os.path.join(out_dir, "spinup_default.npz"),
                        n_solves=len(spin), ic=sim.mData["initial_cond"],
                        **static_pack(sim), **tape_to_arrays(keep))

    _TAPE.clear()
    sim.mData["dim_t"] = n_windows + 1     # run() loops range(1, dim_t)
    sim.run()
    run = list(_TAPE)
    print(f"  run: {len(run)} windows, failed={sum(r['status'] != 0 for r in run)}")
    o = sim.output
    np.savez_compressed(os.path.join(out_dir, "windows_default.npz"),
                        ic=sim.mData["initial_cond"], **static_pack(sim),
                        theta_vol=o["theta_vol"], psi_press=o["psi_press"],
                        K_hrc=o["K_hrc"], K_bkg=o["K_bkg"],
                        wtd_est_cm=o["wtd_est_cm"], abs_error=o["abs_error"],
                        out_lateral_flow=o["lateral_flow"],
                        **tape_to_arrays(run))
    return sim


def gen_variant_runs(out_dir, ic_default):
    """Short taped runs for other closures / profiles / soil parameters.
    The IC is the default spun-up state (a fresh spin-up per variant would
    take minutes each); each variant then runs 12 windows."""
    forcing = synthetic_forcing(365)
    variants = {}
    base = default_params(SITE)
    variants["vg_strat"] = default_params(
        SITE, Hydrological_Model={"Name": "vanGenuchten", "Porosity_Profile": "Stratified"})
    for prof in ("Constant", "Linear", "Exponential"):
        variants["vf_" + prof.lower()] = default_params(
            SITE, Hydrological_Model={"Name": "vrettas_fung", "Porosity_Profile": prof})
    for c in (0, 1, 2):
        variants[f"ens{c}"], _ = ensemble_member(base, c, 1024)
    for c in (3, 4, 5, 6):
        variants[f"mix{c}"], _ = ensemble_member(base, c, 1024, mixed=True)
    variants["lam2"] = default_params(
        SITE, Hydraulic_Conductivity={"Lambda_Exponent": 2.0, "Sigma_Noise": 1.0})
    blob = {}
    import json
    for name, params in variants.items():
        _TAPE.clear()
        sim = make_sim(params, forcing, seed=77, ic=ic_default)
        sim.mData["dim_t"] = 13
        sim.run()
        tape = list(_TAPE)
        print(f"  variant {name}: {len(tape)} windows, "
              f"failed={sum(r['status'] != 0 for r in tape)}")
        blob.update({f"{name}/{k}": v for k, v in static_pack(sim).items()})
        blob.update(tape_to_arrays(tape, prefix=f"{name}/"))
        blob[f"{name}/theta_vol"] = sim.output["theta_vol"]
        blob[f"{name}/K_hrc"] = sim.output["K_hrc"]
        blob[f"{name}/K_bkg"] = sim.output["K_bkg"]
        blob[f"{name}/params_json"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(out_dir, "windows_variants.npz"), **blob)


def gen_rhs_cases(out_dir, sim_default):
    """(psi, args) -> dpsi/dt for many states; also closure outputs on nodes."""
    import json
    forcing = synthetic_forcing(365)
    rng = np.random.default_rng(99)
    ic = sim_default.mData["initial_cond"]
    z = sim_default.mData["z_grid"]
    nx = z.size
    base = default_params(SITE)
    cfgs = {"default": base,
            "vg": default_params(SITE, Hydrological_Model={
                "Name": "vanGenuchten", "Porosity_Profile": "Stratified"}),
            "lin": default_params(SITE, Hydrological_Model={
                "Name": "vrettas_fung", "Porosity_Profile": "Linear"}),
            "exp_nolf": default_params(
                SITE, Hydrological_Model={"Name": "vrettas_fung",
                                          "Porosity_Profile": "Exponential"},
                Simulation_Flags={"SPINUP": False, "ET": True, "LF": False,
                                  "HLIFT": False, "PREDICT": False}),
            "ens7": ensemble_member(base, 7, 1024)[0],
            "ens8": ensemble_member(base, 8, 1024)[0],
            "mix9": ensemble_member(base, 9, 1024, mixed=True)[0],
            "lam2": default_params(SITE, Hydraulic_Conductivity={
                "Lambda_Exponent": 2.0, "Sigma_Noise": 1.0})}
    # states: saturated IC of the spin-up, the spun-up state, a drier state,
    # a state with a perched saturated lens, random rough states.
    sat_ic = np.arange(nx) * 5.0
    states = {"spun": ic, "sat": sat_ic,
              "dry": ic - 50.0 - 30.0 * rng.random(nx),
              "rough": ic + 20.0 * rng.standard_normal(nx),
              "wet_top": np.where(z < 60.0, 2.0 + 0.01 * z, ic)}
    blob = {}
    for cname, params in cfgs.items():
        sim = make_sim(params, forcing, seed=5, ic=ic)
        pde, hm = sim.pde_model, sim.mData["hydro_model"]
        blob.update({f"{cname}/{k}": v for k, v in static_pack(sim).items()})
        blob[f"{cname}/params_json"] = np.array(json.dumps(params))
        k = 0
        for sname, psi in states.items():
            for precip in (0.0, 0.3, 7.5):
                for wtd in (40, 170, 399):
                    for spin in (False, True):
                        if (k % 3) and not (precip == 7.5 and wtd == 170):
                            k += 1
                            continue        # thin the full factorial
                        n_rnd = rng.standard_normal(nx)
                        args = {"wtd": wtd, "n_rnd": n_rnd.copy(), "atm": 0.01,
                                "time": pd.Timestamp("2008-10-01 00:00:00"),
                                "interception": 0.1, "precipitation": precip}
                        sim.mData["sim_flags"]["SPINUP"] = spin
                        dydt = pde(0.0, psi.copy(), args)
                        lat = pde.arg_out["lateral_flow"]
                        th, K, C, kb, qinf = hm(psi.copy(), z, args)
                        key = f"{cname}/case{k:03d}"
                        blob[key + "/psi"] = psi
                        blob[key + "/n_rnd"] = n_rnd
                        blob[key + "/scal"] = np.array([precip, wtd, float(spin)])
                        blob[key + "/dydt"] = np.asarray(dydt, dtype=float)
                        blob[key + "/lat"] = np.array(float(lat))
                        blob[key + "/nodes"] = np.stack([th, K, C, kb])
                        blob[key + "/qinf"] = np.array(float(np.ravel(qinf)[0]))
                        k += 1
        sim.mData["sim_flags"]["SPINUP"] = False
        print(f"  rhs cfg {cname}: {sum(1 for x in blob if x.startswith(cname) and x.endswith('/dydt'))} cases")
    np.savez_compressed(os.path.join(out_dir, "rhs_cases.npz"), **blob)


def gen_jac_cases(out_dir, sim_default):
    """scipy's grouped forward-difference Jacobian at two states."""
    pde = sim_default.pde_model
    z = sim_default.mData["z_grid"]
    nx = z.size
    rng = np.random.default_rng(3)
    ic = sim_default.mData["initial_cond"]
    ones, off = np.ones(nx), np.ones(nx - 1)
    structure = sparse.diags((off, ones, off), offsets=(-1, 0, 1)).tocsc()
    groups = group_columns(structure)
    blob = {"groups": groups}
    for name, psi, precip in (("spun", ic, 0.0), ("rain", ic - 5.0, 2.4)):
        args = {"wtd": 170, "n_rnd": rng.standard_normal(nx), "atm": 0.0,
                "time": pd.Timestamp("2008-10-01"), "interception": 0.1,
                "precipitation": precip}

        def fun_vec(t, Y):
            F = np.empty_like(Y)
            for i, yi in enumerate(Y.T):
                F[:, i] = pde(t, yi, args)
            return F
        f = pde(0.0, psi, args)
        J, fac = num_jac(fun_vec, 0.0, psi, np.asarray(f, float), 1e-3, None,
                         (structure, groups))
        J2, fac2 = num_jac(fun_vec, 0.0, psi, np.asarray(f, float), 1e-3, fac,
                           (structure, groups))
        J = J.toarray()
        J2 = J2.toarray()
        blob[name + "/psi"] = psi
        blob[name + "/n_rnd"] = args["n_rnd"]
        blob[name + "/precip"] = np.array(precip)
        blob[name + "/f"] = np.asarray(f, float)
        blob[name + "/dl"] = np.diag(J, -1)
        blob[name + "/d"] = np.diag(J)
        blob[name + "/du"] = np.diag(J, 1)
        blob[name + "/factor"] = fac
        blob[name + "/d_second"] = np.diag(J2)
        blob[name + "/factor_second"] = fac2
    np.savez_compressed(os.path.join(out_dir, "jac_cases.npz"), **blob)
    # group colouring for the fine grid as well (config 5)
    for n in (1001,):
        st = sparse.diags((np.ones(n - 1), np.ones(n), np.ones(n - 1)),
                          offsets=(-1, 0, 1)).tocsc()
        np.save(os.path.join(out_dir, f"groups_{n}.npy"), group_columns(st))


def gen_fine_grid(out_dir):
    """Config 5 flavour: nx = 1001, stiff forcing, a few windows from the
    saturated initial state with SPINUP flags off."""
    import tempfile
    site = fine_grid_site(os.path.join(tempfile.mkdtemp(), "site.json"))
    forcing = synthetic_forcing(365, stiff=True)
    params = default_params(site)
    z = np.arange(0.0, 5000.0 + 5.0, 5.0)
    ic = np.where(z > 800.0, z - 800.0, -0.35 * (800.0 - z) - 1.0)
    _TAPE.clear()
    sim = make_sim(params, forcing, seed=11, ic=ic)
    sim.mData["dim_t"] = 7
    sim.run()
    tape = list(_TAPE)
    print(f"  fine grid: {len(tape)} windows, failed={sum(r['status'] != 0 for r in tape)},"
          f" nfev={[r['nfev'] for r in tape]}")
    np.savez_compressed(os.path.join(out_dir, "windows_fine.npz"), ic=ic,
                        **static_pack(sim), **tape_to_arrays(tape))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--windows", type=int, default=150)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    print("default run ...")
    sim = gen_default_run(a.out, a.windows)
    print("rhs cases ...")
    gen_rhs_cases(a.out, sim)
    print("jacobian cases ...")
    gen_jac_cases(a.out, sim)
    print("variant runs ...")
    gen_variant_runs(a.out, sim.mData["initial_cond"])
    print("fine grid ...")
    gen_fine_grid(a.out)


if __name__ == "__main__":
    main()
# (tests/golden/make_golden_et.py imports this module for its harness)
