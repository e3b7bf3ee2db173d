"""
CPU tests: the oracle (oracle/hb_oracle.c) against the golden outputs of the
unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py)
and against the reference's own known-answer vectors.
"""
import numpy as np
import pytest

import oracle
from hydromodel_b200 import pack
from helpers import (column_from_json, default_column, rhs_cases, rhs_configs, tol_scaled)


def test_find_wtd_reference_vectors():
    """The four exact vectors of the reference's own test
    (code/tests/test_utilities.py:56-86): expected 9, 0, 6, 7."""
    from hydromodel_b200.src.utilities import find_wtd
    cases = (([0] * 10, 9), ([1] * 10, 0), ([0, 0, 0, 0, 0, 0, 1, 1, 1, 1], 6),
             ([0, 1, 0, 1, 0, 0, 0, 1, 1, 1], 7))
    for vec, want in cases:
        v = np.array(vec, dtype=bool)
        assert oracle.find_wtd(v) == want
        assert find_wtd(v) == want


def test_fd_colouring_matches_scipy(golden):
    g = golden("jac_cases.npz")
    grp, ng = pack.fd_groups(401)
    assert ng == 5 and np.array_equal(grp, g["groups"])


def test_fd_colouring_1001(golden):
    import os
    from conftest import GOLDEN
    grp, ng = pack.fd_groups(1001)
    assert ng == 5 and np.array_equal(grp, np.load(os.path.join(GOLDEN, "groups_1001.npy")))


def test_static_pack_matches_reference_profiles(golden):
    g = golden("rhs_cases.npz")
    for cfg in rhs_configs(g):
        col = column_from_json(g[f"{cfg}/params_json"])
        assert np.array_equal(col["x"], g[f"{cfg}/z"])
        assert np.array_equal(col["objs"]["porosity"].profile, g[f"{cfg}/porosity"])
        assert np.array_equal(col["forcing"]["zWtd_cm"], g[f"{cfg}/zwtd_cm"])
        assert np.array_equal(col["forcing"]["atm"], g[f"{cfg}/atm_series"])
        assert col["forcing"]["surface_evap"] == float(g[f"{cfg}/surface_evap"])


def test_rhs_against_reference(golden):
    """RichardsPDE.__call__ parity.  Tolerance: 1e-12 of the magnitude sum of
    the terms whose (cancelling) sum forms dpsi/dt; closure outputs 1e-12
    relative (measured: 6e-14 and 5e-14)."""
    g = golden("rhs_cases.npz")
    n_cases = 0
    for cfg in rhs_configs(g):
        col = column_from_json(g[f"{cfg}/params_json"])
        for _, psi, nr, precip, wtd, spin, p in rhs_cases(g, cfg):
            dydt, lat, cond = oracle.rhs(col["x"], col["par"], col["stat"], psi, nr, precip,
                                         wtd, 0, spin, with_cond=True)
            ref = g[p + "dydt"]
            assert np.all(np.abs(dydt - ref) <= 1e-12 * cond + 1e-300), (cfg, p)
            nodes, qinf = oracle.closure_nodes(col["x"], col["par"], col["stat"], psi, nr)
            assert np.allclose(nodes, g[p + "nodes"], rtol=1e-12, atol=0), (cfg, p)
            assert qinf == pytest.approx(float(g[p + "qinf"]), rel=1e-12)
            assert lat == pytest.approx(float(g[p + "lat"]), rel=1e-12, abs=1e-300)
            n_cases += 1
    assert n_cases >= 250


def test_num_jac_against_scipy(golden):
    """Tridiagonal FD Jacobian: entries within 1e-5 relative (forward
    differences amplify 1e-16 differences of f by 1/h ~ 1e8), zero pattern and
    step factors identical."""
    g = golden("jac_cases.npz")
    col = default_column()
    grp, _ = pack.fd_groups(401)
    for name in ("spun", "rain"):
        psi, nr, precip = g[name + "/psi"], g[name + "/n_rnd"], float(g[name + "/precip"])
        dl, d, du, fac = oracle.num_jac(col["x"], grp, col["par"], col["stat"], psi, nr, precip, 170)
        for mine, ref in ((dl, np.r_[0, g[name + "/dl"]]), (d, g[name + "/d"]),
                          (du, np.r_[g[name + "/du"], 0])):
            assert np.array_equal(mine == 0, ref == 0)
            assert np.allclose(mine, ref, rtol=1e-5, atol=0)
        assert np.array_equal(fac, g[name + "/factor"])
        _, d2, _, fac2 = oracle.num_jac(col["x"], grp, col["par"], col["stat"], psi, nr, precip,
                                        170, fac=fac)
        assert np.array_equal(fac2, g[name + "/factor_second"])
        assert np.allclose(d2, g[name + "/d_second"], rtol=1e-5, atol=0)


def _teacher_forced(g, prefix, col, spinup):
    grp, _ = pack.fd_groups(col["x"].size)
    n = g[prefix + "y0"].shape[0]
    errs, same, th_err, wtd_ok = [], [], [], []
    for w in range(n):
        nr = g[prefix + "n_rnd"][w].copy()
        t0 = float(g[prefix + "t0"][w])
        y, st, status, lat, t_acc = oracle.solve_window(
            col["x"], grp, col["par"], col["stat"], t0, t0 + 1.0, g[prefix + "y0"][w], nr,
            float(g[prefix + "precip"][w]), int(g[prefix + "wtd"][w]), 0, spinup)
        ref = g[prefix + "y_end"][w]
        assert status == int(g[prefix + "status"][w] != 0)
        errs.append(tol_scaled(y, ref))
        same.append(st["nfev"] == g[prefix + "nfev"][w] and st["njev"] == g[prefix + "njev"][w]
                    and st["nlu"] == g[prefix + "nlu"][w] and st["nsteps"] == g[prefix + "nsteps"][w])
        psi_sat = col["par"][pack.P_PSI_SAT]
        wtd_ok.append(oracle.find_wtd(y >= psi_sat) == oracle.find_wtd(ref >= psi_sat))
        th = oracle.closure_nodes(col["x"], col["par"], col["stat"], y, nr)[0][0]
        th_ref = oracle.closure_nodes(col["x"], col["par"], col["stat"], ref, nr)[0][0]
        th_err.append(float(np.max(np.abs(th - th_ref) / th_ref)))
    return np.array(errs), np.array(same), np.array(th_err), np.array(wtd_ok)


def test_windows_default_run_against_reference(golden):
    """Teacher-forced RichardsPDE.solve parity on 150 windows of the default
    run.  The integrator is sensitive to rounding (the FD Jacobian amplifies
    1e-16 differences), so parity is stated in units of the
    solver's own tolerance atol + rtol*|psi|:
      * solver counters (nfev, njev, nlu, steps) identical in >= 95 % of windows
      * psi within 0.1 tolerance units in >= 95 % of windows, always < 5
      * theta relative error <= 1e-6 in >= 98 % of windows
      * water-table node identical in every window."""
    g = golden("windows_default.npz")
    errs, same, th_err, wtd_ok = _teacher_forced(g, "", default_column(), 0)
    assert same.mean() >= 0.95
    assert np.mean(errs <= 0.1) >= 0.95 and errs.max() < 5.0
    assert np.median(errs) < 1e-4
    assert np.mean(th_err <= 1e-6) >= 0.98
    assert wtd_ok.all()


def test_windows_spinup_against_reference(golden):
    """Spin-up windows (SPINUP flags) start from the saturated column and are
    the stiffest: decision flips are more frequent there."""
    g = golden("spinup_default.npz")
    errs, same, th_err, wtd_ok = _teacher_forced(g, "", default_column(), 1)
    assert same.mean() >= 0.75
    assert np.median(errs) < 1e-2
    assert wtd_ok.mean() >= 0.9


@pytest.mark.parametrize("name", ["vg_strat", "vf_constant", "vf_linear", "vf_exponential",
                                  "ens0", "ens1", "ens2", "mix3", "mix4", "mix5", "mix6", "lam2"])
def test_windows_variants_against_reference(golden, name):
    g = golden("windows_variants.npz")
    col = column_from_json(g[f"{name}/params_json"])
    errs, same, th_err, wtd_ok = _teacher_forced(g, name + "/", col, 0)
    assert same.mean() >= 0.9
    assert np.mean(errs <= 0.2) >= 0.9 and errs.max() < 5.0
    assert wtd_ok.all()


def test_windows_fine_grid_against_reference(golden):
    g = golden("windows_fine.npz")
    from hydromodel_b200.synthetic import default_params, synthetic_forcing, fine_grid_site
    import tempfile, os
    site = fine_grid_site(os.path.join(tempfile.mkdtemp(), "site.json"))
    col = pack.pack_column(default_params(site), synthetic_forcing(365, stiff=True))
    assert col["x"].size == 1001
    errs, same, th_err, wtd_ok = _teacher_forced(g, "", col, 0)
    assert same.mean() >= 0.6
    assert np.mean(errs <= 0.5) >= 0.8


def test_run_driver_matches_window_by_window(golden):
    """orc_run_c (the free-running window loop with the noise-refresh rule of
    simulation.py:599-602) reproduces the teacher-forced chain of solves."""
    g = golden("windows_default.npz")
    col = default_column()
    grp, _ = pack.fd_groups(401)
    T = 31
    f = col["forcing"]
    # rebuild the tape from the golden run: slot 0 = first draw, then refreshes
    tape = [g["n_rnd"][0] if not (f["precipitation_cm"][1] > 0.5) else None]
    base = None
    slots = []
    for i in range(1, T):
        refreshed = f["precipitation_cm"][i] > 0.5 or i % 48 == 0
        if refreshed:
            slots.append(g["n_rnd"][i - 1])
        elif base is None:
            base = g["n_rnd"][i - 1]
    tape = np.stack([base] + slots)
    psi, wscal, fields, stats, used = oracle.run(
        col["x"], grp, col["par"], col["stat"], f["precipitation_cm"][:T], f["wtd_idx"][:T],
        f["daylight"][:T], f["atm"][:T], tape, g["ic"], out_stride=10)
    assert used == tape.shape[0]
    assert tol_scaled(psi, g["psi_press"][T - 1]) < 1.0
    assert np.array_equal(wscal[:, 0] * 5.0, g["wtd_est_cm"][:T])
    assert fields.shape[0] == 4
    assert np.allclose(fields[1, 1], g["theta_vol"][10], rtol=1e-5)


def _et_cases(g, cfg):
    names = sorted({k.split("/")[1] for k in g.files if k.startswith(cfg + "/case")})
    for cs in names:
        p = f"{cfg}/{cs}/"
        precip, wtd, hour, atm = g[p + "scal"]
        yield p, g[p + "psi"], g[p + "n_rnd"], float(precip), int(wtd), int(6 <= hour < 18), float(atm)


def test_et_and_hlift_sinks_against_reference(golden):
    """Root-water uptake (daylight) and hydraulic redistribution (night):
    RichardsPDE.__call__ with the reference's tree-root efficiency
    (richards_pde.py:225-303, tree_roots.py:179-292), 200 cases over 5 configs."""
    g = golden("et_cases.npz")
    n = n_tr = 0
    for cfg in sorted({k.split("/")[0] for k in g.files}):
        col = column_from_json(g[f"{cfg}/params_json"])
        for p, psi, nr, precip, wtd, day, atm in _et_cases(g, cfg):
            dydt, lat, cond, tr = oracle.rhs(col["x"], col["par"], col["stat"], psi, nr, precip, wtd,
                                             day, 0, atm=atm, with_transpiration=True)
            assert np.all(np.abs(dydt - g[p + "dydt"]) <= 1e-12 * cond + 1e-300), p
            ref_tr = float(g[p + "transp"])
            assert tr == pytest.approx(ref_tr, rel=1e-12, abs=1e-300)
            n += 1; n_tr += ref_tr != 0.0
    assert n == 200 and n_tr >= 60


def test_halfhourly_daylight_windows_against_reference_tape(golden):
    """Reference-taped half-hourly run with the default flags (ET on) and a gentle demand
    (tests/golden/make_golden_et.py --halfhourly-gentle): 1 night + 23 daylight windows with the
    root-water-uptake sink active, teacher-forced.  Transpiration of the last RHS evaluation
    (arg_out, simulation.py:630) must equal the tape's; end states as in the daily windows."""
    from helpers import column_from_json
    g = golden("windows_halfhourly_gentle.npz")
    col = column_from_json(g["params_json"])
    grp, _ = pack.fd_groups(col["x"].size)
    errs, same = [], []
    n_day = 0
    for w in range(g["y0"].shape[0]):
        nr = g["n_rnd"][w].copy()
        day = int(6 <= g["hour"][w] < 18)
        t0 = float(g["t0"][w])
        y, st, status, lat, t_acc = oracle.solve_window(
            col["x"], grp, col["par"], col["stat"], t0, t0 + 1.0, g["y0"][w], nr, float(g["precip"][w]),
            int(g["wtd"][w]), day, 0, atm=float(g["atm"][w]))
        assert status == 0 and g["status"][w] == 0
        errs.append(tol_scaled(y, g["y_end"][w]))
        same.append(st["nfev"] == g["nfev"][w] and st["njev"] == g["njev"][w] and st["nlu"] == g["nlu"][w]
                    and st["nsteps"] == g["nsteps"][w])
        n_day += day
    errs, same = np.array(errs), np.array(same)
    print("oracle half-hourly windows: %d daylight; counters-match %.3f; err median %.2e max %.2e"
          % (n_day, same.mean(), np.median(errs), errs.max()))
    assert n_day >= 20
    assert same.mean() >= 0.85
    assert np.median(errs) < 1e-3 and errs.max() < 5.0


def test_one_pass_jacobian_differences_equal_the_evaluation_per_group(golden):
    """The identity the device's jac_fused rests on, checked on the CPU: the rows j-1..j+1 of
    f(y + h_g) that num_jac reads for column j (SCIPY common.py:385-443) are the same numbers,
    bit for bit, whether f is evaluated once per colour group or assembled from two closure
    evaluations per midpoint (own node perturbed / right neighbour perturbed) with the water table
    of the lateral-flow sink estimated per group.  States: the spun-up profile and states of the
    golden run, with perturbations from num_jac's own step rule up to ones that move cells across
    saturation at the water table; flags: lateral flow on / off, hydraulic lift at night, surface
    evaporation by day (root uptake off), spin-up; both closures; scipy's 5 colour groups and the
    3 groups j mod 3."""
    from hydromodel_b200.synthetic import default_params, ensemble_member
    from helpers import SITE, forcing365
    g = golden("windows_default.npz")
    rng = np.random.default_rng(2024)
    base = default_params(SITE)
    table = forcing365()
    cols = [pack.pack_column(base, table)]
    for c in (3, 11):
        p, _ = ensemble_member(base, c, 1024, mixed=True)
        cols.append(pack.pack_column(p, table))
    n_checked = 0
    for ci, col in enumerate(cols):
        nx = col["x"].size
        for ng_variant in ("scipy", "mod3"):
            grp = pack.fd_groups(nx)[0] if ng_variant == "scipy" else (np.arange(nx) % 3).astype(np.int32)
            for w in (0, 40, 90):
                y = g["y0"][w].copy() if nx == g["y0"].shape[1] else None
                if y is None:
                    continue
                nr = g["n_rnd"][w]
                for scale in (1.4901161193847656e-08, 1e-5, 3e-2, -4.0):
                    sign = np.where(rng.random(nx) < 0.5, 1.0, -1.0)
                    # (a negative scale is an absolute step in cm: moves midpoints across psi_sat)
                    h = sign * (scale * np.maximum(1e-3, np.abs(y)) if scale > 0 else -scale)
                    for flags, day, spin in ((1.0, 0, 0), (0.0, 0, 0), (5.0, 0, 0), (1.0, 1, 0), (1.0, 0, 1)):
                        par = col["par"].copy()
                        par[pack.P_FLAGS] = flags
                        kw = dict(daylight=day, spinup=spin)
                        a = oracle.group_rows(col["x"], grp, par, col["stat"], y, nr, 0.7, 170, h, fused=False, **kw)
                        b = oracle.group_rows(col["x"], grp, par, col["stat"], y, nr, 0.7, 170, h, fused=True, **kw)
                        assert np.array_equal(a.view(np.int64), b.view(np.int64)), (ci, ng_variant, w, scale, flags, day, spin)
                        n_checked += 1
    assert n_checked >= 250
    # the absolute steps do move midpoints across saturation next to the water table
    y = g["y0"][40]
    psi_sat = cols[0]["par"][pack.P_PSI_SAT]
    mid = 0.5 * (y[:-1] + y[1:])
    assert np.sum(np.abs(mid - psi_sat) < 2.0) >= 1
    # root uptake active: refused, as on the device (one evaluation per group there)
    par = cols[0]["par"].copy(); par[pack.P_FLAGS] = 3.0
    with pytest.raises(ValueError):
        oracle.group_rows(cols[0]["x"], pack.fd_groups(401)[0], par, cols[0]["stat"], y, g["n_rnd"][40], 0.7, 170,
                          1e-8 * np.abs(y), daylight=1, fused=True)
