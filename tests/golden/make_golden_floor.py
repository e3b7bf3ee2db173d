#!/usr/bin/env python3
"""
Noise floor of the UNMODIFIED reference against itself (tests/golden/noise_floor.npz).

The reference's integrator is not reproducible beyond a floor: its Jacobian is a
forward difference (rounding amplified by 1/h ~ 1e8) and its Newton iteration stops
at 0.03 in scaled norm, so a relative perturbation of 1e-14 of a window's start
state moves the end state -- by rounding level while the solver takes the same
decisions, by O(1e-3..1) solver-tolerance units once one step-size / convergence
decision flips.  Every tolerance of the device parity tests is derived from the
arrays recorded here (tests/test_gpu_parity.py), instead of being stated by hand.

Recorded (reference @ container stack, see make_golden.py):
  tf_*      teacher-forced twin solves of the 150 windows of the default run
            (simulation.py:576-649): for each window the reference's own solve and
            a second solve_ivp call from y0 * (1 + eps), eps = 1e-14: end-state
            difference in solver-tolerance units, theta relative difference,
            whether the four counters agree (`same`), whether in addition the
            accepted time grids coincide (`grid`: same decisions), water-table node
            difference; `*_e12`: the same for four twins that model an independent
            implementation: start state AND every RHS evaluation perturbed by random
            patterns of relative size 1e-12 (the RHS parity bar of device and oracle).
  spin_*    the same for the first 24 and last 8 solves of the spin-up
            (simulation.py:444-480) -- the saturated start is the stiffest regime.
  free_*    free-running twins: two reference runs of 150 windows from IC and
            IC * (1 + eps) with the same seed (no spin-up, same noise draws), for
            eps = 1e-14 and 1e-12: per-row psi difference in tolerance units,
            theta relative difference, water-table node difference.

Usage:  python tests/golden/make_golden_floor.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                           # noqa: E402
from hydromodel_b200.synthetic import synthetic_forcing, default_params   # noqa: E402

EPS = 1.0e-14
_TWIN = []
_real = mg._real_solve_ivp


def _units(y, ref):
    return float(np.max(np.abs(y - ref) / (1e-3 + 1e-3 * np.abs(ref))))


_RNG = np.random.default_rng(4321)
N_E12 = 4          # random 1e-12 perturbation patterns per solve


def _same(a, b):
    """Same decisions: the four counters agree AND the accepted time grids coincide."""
    counters = (a.nfev, a.njev, a.nlu, a.t.size) == (b.nfev, b.njev, b.nlu, b.t.size)
    grid = counters and bool(np.allclose(a.t, b.t, rtol=1e-6, atol=1e-9))
    return int(counters), int(grid)


def _twin_solve_ivp(fun, t_span, y0, **kw):
    """The reference's call, preceded by the same call from perturbed starts: y0 * (1 + 1e-14)
    and N_E12 random patterns y0 * (1 + 1e-12 * N(0,1)) -- the size of the device's and the
    oracle's rounding differences in the right-hand side (1e-12 relative, RHS parity test)."""
    y0 = np.asarray(y0, dtype=float)
    twin = _real(fun, t_span=t_span, y0=y0 * (1.0 + EPS), **kw)
    # an independent implementation differs in the rounding of every RHS evaluation, not only in
    # the start state: the 1e-12 twins also multiply dpsi/dt by (1 + 1e-12 * pattern), a fixed
    # random pattern per twin (the RHS parity bar of the device and of the oracle is 1e-12)
    twins12 = []
    for _ in range(N_E12):
        pat_y, pat_f = _RNG.standard_normal(y0.size), _RNG.standard_normal(y0.size)

        def fun_pert(t, y, *a, _p=pat_f):
            return np.asarray(fun(t, y, *a)) * (1.0 + 1e-12 * _p)
        twins12.append(_real(fun_pert, t_span=t_span, y0=y0 * (1.0 + 1e-12 * pat_y), **kw))
    sol = _real(fun, t_span=t_span, y0=y0, **kw)
    c14, g14 = _same(twin, sol)
    rec = {"units": _units(twin.y[:, -1], sol.y[:, -1]), "same": c14, "grid": g14,
           "y_end": sol.y[:, -1].copy(), "y_twin": twin.y[:, -1].copy(),
           "units12": [_units(t.y[:, -1], sol.y[:, -1]) for t in twins12],
           "same12": [_same(t, sol)[0] for t in twins12], "grid12": [_same(t, sol)[1] for t in twins12]}
    _TWIN.append(rec)
    return sol


def _theta_rel(sim, y_a, y_b, n_rnd):
    hm, z = sim.mData["hydro_model"], sim.mData["z_grid"]
    th_a = hm(y_a, z, {"n_rnd": n_rnd})[0]
    th_b = hm(y_b, z, {"n_rnd": n_rnd})[0]
    return float(np.max(np.abs(th_a - th_b) / np.abs(th_a)))


def _wtd(sim, y):
    from src.utilities import find_wtd
    return int(find_wtd(y >= sim.mData["soil"].psi_sat))


def main(n_windows=150):
    forcing = synthetic_forcing(365)
    params = default_params(mg.SITE)
    blob = {"eps": np.array(EPS)}
    mg.ref_pde.solve_ivp = _twin_solve_ivp

    # ---- spin-up and run of the default golden simulation, every solve twinned ----
    _TWIN.clear()
    sim = mg.make_sim(params, forcing, seed=1234)
    spin = list(_TWIN)
    keep = spin[:24] + spin[-8:]
    zeros = np.zeros(sim.mData["z_grid"].size)
    blob["spin_units"] = np.array([r["units"] for r in keep])
    blob["spin_same"] = np.array([r["same"] for r in keep])
    blob["spin_grid"] = np.array([r["grid"] for r in keep])
    blob["spin_units_e12"] = np.array([r["units12"] for r in keep])
    blob["spin_same_e12"] = np.array([r["same12"] for r in keep])
    blob["spin_grid_e12"] = np.array([r["grid12"] for r in keep])
    blob["spin_wtd_diff"] = np.array([abs(_wtd(sim, r["y_end"]) - _wtd(sim, r["y_twin"])) for r in keep])
    blob["spin_n_solves"] = np.array(len(spin))
    print(f"  spin-up twins: {len(spin)} solves; kept units {blob['spin_units'].round(3)}")
    _TWIN.clear()
    sim.mData["dim_t"] = n_windows + 1
    sim.run()
    tf = list(_TWIN)
    blob["tf_units"] = np.array([r["units"] for r in tf])
    blob["tf_same"] = np.array([r["same"] for r in tf])
    blob["tf_grid"] = np.array([r["grid"] for r in tf])
    blob["tf_units_e12"] = np.array([r["units12"] for r in tf])
    blob["tf_same_e12"] = np.array([r["same12"] for r in tf])
    blob["tf_grid_e12"] = np.array([r["grid12"] for r in tf])
    blob["tf_theta_rel"] = np.array([_theta_rel(sim, r["y_end"], r["y_twin"], zeros) for r in tf])
    blob["tf_wtd_diff"] = np.array([abs(_wtd(sim, r["y_end"]) - _wtd(sim, r["y_twin"])) for r in tf])
    ic = np.array(sim.mData["initial_cond"], dtype=float)
    print("  teacher-forced twins: same counters %.3f; units max (same) %.2e, max (flipped) %.2e" %
          (blob["tf_same"].mean(), blob["tf_units"][blob["tf_same"] == 1].max(),
           blob["tf_units"][blob["tf_same"] == 0].max() if (blob["tf_same"] == 0).any() else 0.0))

    # ---- free-running twins ---------------------------------------------------------
    mg.ref_pde.solve_ivp = _real
    runs = {}
    for tag, eps in (("base", 0.0), ("e14", 1e-14), ("e12", 1e-12)):
        s = mg.make_sim(params, forcing, seed=1234, ic=ic * (1.0 + eps))
        s.mData["dim_t"] = n_windows + 1
        s.run()
        runs[tag] = (s, s.output["psi_press"].copy(), s.output["theta_vol"].copy(),
                     s.output["wtd_est_cm"].copy())
    for tag in ("e14", "e12"):
        _, psi, th, wtd = runs[tag]
        _, psi0, th0, wtd0 = runs["base"]
        blob[f"free_{tag}_units"] = np.array([_units(psi[i], psi0[i]) for i in range(n_windows + 1)])
        blob[f"free_{tag}_theta_rel"] = np.max(np.abs(th - th0) / np.abs(th0), axis=1)
        blob[f"free_{tag}_wtd_diff"] = np.abs(wtd - wtd0) / 5.0
        print(f"  free-running {tag}: psi units max {blob[f'free_{tag}_units'].max():.3f}, "
              f"theta rel max {blob[f'free_{tag}_theta_rel'].max():.2e}, "
              f"wtd node diff max {blob[f'free_{tag}_wtd_diff'].max():.0f}")
    np.savez_compressed(os.path.join(HERE, "noise_floor.npz"), **blob)


if __name__ == "__main__":
    main()
