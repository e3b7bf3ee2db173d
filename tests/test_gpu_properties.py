"""
Size-independent properties of the device path at the sizes BASELINE.json
quotes (configs[1], [3], [4]; run on the B200 box).  The oracle finishes a
handful of column-years in seconds, not a thousand, so at full size the CUDA
path is checked through what must hold whatever the numbers are:

* determinism      -- the same job twice gives the same bits;
* independence     -- a column's result does not depend on which other columns
                      share the launch, on the order the persistent CTAs pick
                      the work items up, or on how the ensemble is split over
                      contexts (= GPUs): columns are independent units of the
                      reference (one Simulation object each);
* invariants       -- theta in [theta_res, porosity], K <= K_bkg, water-table
                      index inside the grid and tracking the observations the
                      lateral-flow sink steers to (richards_pde.py:305-377),
                      counters consistent with scipy's bookkeeping;
* a sample against the oracle on the same inputs.
"""
import os

import numpy as np
import pytest

import oracle
from hydromodel_b200 import capi, pack
from hydromodel_b200.synthetic import (default_params, ensemble_member, fine_grid_site,
                                       synthetic_forcing)
from helpers import SITE

pytestmark = pytest.mark.gpu


def pack_ensemble(n_cols, forcing, mixed=False, site=SITE, first=0, total=1024):
    base = default_params(site)
    cols = []
    for c in range(n_cols):
        p, mult = ensemble_member(base, first + c, total, mixed=mixed)
        cols.append(pack.pack_column(p, forcing, precip_mult=mult))
    return cols


def tape_for(cols, n_cols, seed=0, first=0):
    f = cols[0]["forcing"]
    rows = np.arange(f["dim_t"])
    mult_max = max(float(c["par"][pack.P_PRECIP_MULT]) for c in cols)
    refresh = ((mult_max * f["precipitation_cm"] > 0.5) | (rows % 48 == 0)) & (rows > 0)
    slots = 1 + int(refresh.sum())
    nx = cols[0]["x"].size
    tape = np.empty((n_cols, slots, nx))
    for c in range(n_cols):
        tape[c] = np.random.default_rng(np.random.SeedSequence(seed + first + c)).standard_normal((slots, nx))
    return tape


def run_job(cols, ic, tape=None, philox=None, out_stride=30, rows=None, sel=None):
    """Run the columns `sel` (default: all) of `cols` in one context; returns a dict of arrays."""
    sel = list(range(len(cols))) if sel is None else list(sel)
    c0 = cols[0]
    nx, f = c0["x"].size, c0["forcing"]
    T = f["dim_t"] if rows is None else rows
    grp, ng = pack.fd_groups(nx)
    ctx = capi.Context(nx, len(sel))
    ctx.set_grid(c0["x"], grp, ng)
    ctx.set_columns(np.stack([cols[k]["par"] for k in sel]), np.stack([cols[k]["stat"] for k in sel]))
    ctx.set_forcing(f["precipitation_cm"][:T], f["wtd_idx"][:T], f["daylight"][:T], f["atm"][:T])
    ctx.set_output(out_stride)
    if tape is not None:
        ctx.set_noise_tape(np.ascontiguousarray(tape[sel]))
    else:
        ctx.set_noise_philox(philox[0], philox[1])
    psi0 = ic if ic.ndim == 2 else np.tile(ic, (len(cols), 1))
    ctx.set_state(np.ascontiguousarray(psi0[sel]))
    ctx.reset_counters()
    ctx.run(0, T)
    out = {"wscal": ctx.get_wscal()[:, :T], "state": ctx.get_state(), "counters": ctx.get_counters()}
    if out_stride > 0:
        out["fields"] = ctx.get_fields()
    ctx.close()
    return out


def same_bits(a, b, sel=None):
    for k in a:
        x = a[k] if sel is None else a[k][sel]
        if not np.array_equal(x, b[k]):
            return k
    return None


def check_invariants(cols, out, T, track_cells=25, track_frac=0.99):
    nx = cols[0]["x"].size
    w = out["wscal"]
    wtd = w[:, :, pack.W_WTD_EST]
    assert np.all(np.isfinite(out["state"])) and np.all(np.isfinite(w))
    assert np.all((wtd >= 0) & (wtd <= nx - 1))
    status = w[:, :, pack.W_STATUS]
    assert np.all(status[:, 1:] == 0), "integration failures on the synthetic forcing"
    cn = out["counters"]
    assert np.all(cn[:, pack.C_NWINDOWS] == T - 1)
    # a failed attempt is retried with damped noise (quirk Q4); the row status stays 0
    assert np.all(cn[:, pack.C_NFAIL] <= 0.02 * cn[:, pack.C_NWINDOWS] + 1)
    # scipy bookkeeping: every accepted or rejected step costs at least one Newton
    # evaluation; the FD Jacobian costs at most 1 + (groups + retries) evaluations;
    # LU factorisations cannot outnumber the Newton solves they serve
    assert np.all(cn[:, pack.C_NFEV] >= cn[:, pack.C_NSTEPS] + cn[:, pack.C_NREJECT])
    assert np.all(cn[:, pack.C_NRHS] >= cn[:, pack.C_NFEV] + 5 * cn[:, pack.C_NJEV])
    assert np.all(cn[:, pack.C_NLU] <= cn[:, pack.C_NNEWTON])
    assert np.all(cn[:, pack.C_NJEV] >= cn[:, pack.C_NWINDOWS])
    # the monitoring sink steers the water table towards the observations (oracle on
    # this forcing: never more than 14 cells away)
    obs = cols[0]["forcing"]["wtd_idx"][:T]
    ok = np.abs(wtd[:, 1:] - obs[None, 1:]) <= track_cells
    assert ok.mean() >= track_frac, ok.mean()
    if "fields" in out:
        fl = out["fields"]
        th, K, kb = fl[:, :, pack.F_THETA], fl[:, :, pack.F_K], fl[:, :, pack.F_KBKG]
        por = np.stack([c["objs"]["porosity"].profile for c in cols])[:, None, :]
        thr = np.stack([c["par"][pack.P_THETA_RES] for c in cols])[:, None, None]
        assert np.all(th <= por * (1 + 1e-12)) and np.all(th >= thr * (1 - 1e-12))
        assert np.all(K <= kb * (1 + 1e-12)) and np.all(K >= 0.0)


def oracle_sample(cols, tape, ic, out, sample, rows, cells=1, frac=0.95):
    """Free-running oracle on a few columns: water-table index within `cells`
    in >= `frac` of the rows (trajectories diverge at rounding level
    through the finite-difference Jacobian, the diagnostics do not)."""
    c0 = cols[0]
    f = c0["forcing"]
    grp, _ = pack.fd_groups(c0["x"].size)
    for k in sample:
        psi0 = ic[k] if ic.ndim == 2 else ic
        _, wscal, _, stats, _ = oracle.run(c0["x"], grp, cols[k]["par"], cols[k]["stat"],
                                           f["precipitation_cm"][:rows], f["wtd_idx"][:rows],
                                           f["daylight"][:rows], f["atm"][:rows], tape[k], psi0)
        d = np.abs(wscal[:, pack.W_WTD_EST] - out["wscal"][k, :rows, pack.W_WTD_EST])
        assert (d <= cells).mean() >= frac, (k, d.max(), (d <= cells).mean())


@pytest.fixture(scope="module")
def config1(golden):
    forcing = synthetic_forcing(365)
    cols = pack_ensemble(512, forcing)
    ic = golden("windows_default.npz")["ic"]
    tape = tape_for(cols, len(cols))
    out = run_job(cols, ic, tape=tape)
    return cols, ic, tape, out


def test_config1_full_year_invariants(config1):
    """configs[1]: n / a0 / psi_sat ensemble, 365 daily rows, host noise tape."""
    cols, ic, tape, out = config1
    check_invariants(cols, out, 365)


def test_config1_determinism_and_batch_independence(config1):
    cols, ic, tape, out = config1
    again = run_job(cols, ic, tape=tape)
    assert same_bits(out, again) is None
    sel = [3, 17, 18, 64, 255, 256, 300, 411, 511] + list(range(100, 137))     # ragged subset
    alone = run_job(cols, ic, tape=tape, sel=sel)
    assert same_bits(out, alone, sel=sel) is None


def test_config1_work_item_order_independence(config1, monkeypatch):
    """Time slicing: the chunk length of the work items must not change a bit."""
    cols, ic, tape, out = config1
    sel = list(range(0, 96))
    for chunk in ("1", "7", "400"):
        monkeypatch.setenv("HB_CHUNK", chunk)
        res = run_job(cols, ic, tape=tape, sel=sel)
        assert same_bits(out, res, sel=sel) is None, chunk


def test_compile_time_and_run_time_layout_instances_agree(config1, monkeypatch):
    """401-node columns run the instance compiled for the padded vector length 416;
    the run-time-length instance of the same code must give the same bits."""
    cols, ic, tape, out = config1
    sel = list(range(200, 264))
    monkeypatch.setenv("HB_NO_FIXED", "1")
    res = run_job(cols, ic, tape=tape, sel=sel)
    assert same_bits(out, res, sel=sel) is None


def test_config1_sample_against_oracle(config1):
    cols, ic, tape, out = config1
    oracle_sample(cols, tape, ic, out, sample=[0, 201, 510], rows=60)


def test_philox_sharding_invariance(golden):
    """Device-generated noise depends on (seed, GLOBAL column, row) only: a job
    split over two contexts (= two GPUs) equals the unsplit job bit for bit."""
    forcing = synthetic_forcing(120)
    cols = pack_ensemble(64, forcing)
    ic = golden("windows_default.npz")["ic"]
    whole = run_job(cols, ic, philox=(99, 0), out_stride=40)
    lo = run_job(cols[:40], ic, philox=(99, 0), out_stride=40)
    hi = run_job(cols[40:], ic, philox=(99, 40), out_stride=40)
    for k in whole:
        assert np.array_equal(whole[k][:40], lo[k]) and np.array_equal(whole[k][40:], hi[k]), k
    other = run_job(cols[:8], ic, philox=(100, 0), out_stride=40)
    assert not np.array_equal(other["state"], whole["state"][:8])
    check_invariants(cols, whole, 120)


def test_config3_mixed_models_and_profiles(golden):
    """configs[3]: conductivity model and porosity profile drawn per column."""
    forcing = synthetic_forcing(120)
    cols = pack_ensemble(96, forcing, mixed=True)
    models = {float(c["par"][pack.P_MODEL]) for c in cols}
    assert len(models) == 2, "the draw must contain both conductivity models"
    ic = golden("windows_default.npz")["ic"]
    tape = tape_for(cols, len(cols))
    out = run_job(cols, ic, tape=tape, out_stride=40)
    # (all columns start from the default column's IC, not from their own spin-up:
    # no tracking bar here)
    check_invariants(cols, out, 120, track_frac=0.0)
    for k in (0, 31, 95):                                # each column alone = inside the batch
        alone = run_job(cols, ic, tape=tape, out_stride=40, sel=[k])
        assert same_bits(out, alone, sel=[k]) is None, k
    vg = [k for k, c in enumerate(cols) if c["par"][pack.P_MODEL] != 0.0][:2]
    vf = [k for k, c in enumerate(cols) if c["par"][pack.P_MODEL] == 0.0][:2]
    oracle_sample(cols, tape, ic, out, sample=vg + vf, rows=40)


def test_config4_fine_grid_stiff_forcing(tmp_path):
    """configs[4]: >= 1,000 layers (nx = 1001 at the reference's dz = 5 cm),
    heavier and more frequent rain; the 512-thread kernel variants."""
    site = fine_grid_site(str(tmp_path / "site_fine.json"))
    forcing = synthetic_forcing(40, stiff=True)
    cols = pack_ensemble(24, forcing, site=site)
    nx = cols[0]["x"].size
    assert nx >= 1001
    # start from the saturated profile the reference's spin-up starts from, spun up
    # on the device for a few windows (the spin-up kernel is part of the path)
    grp, ng = pack.fd_groups(nx)
    f = cols[0]["forcing"]
    ctx = capi.Context(nx, len(cols))
    ctx.set_grid(cols[0]["x"], grp, ng)
    ctx.set_columns(np.stack([c["par"] for c in cols]), np.stack([c["stat"] for c in cols]))
    ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])
    ctx.set_state(np.stack([pack.saturated_initial_state(c["objs"]) for c in cols]))
    nr0 = np.stack([np.random.default_rng(7 + k).standard_normal(nx) for k in range(len(cols))])
    iters = ctx.spinup(np.full(len(cols), f["zWtd_cm"][0]), nr0, max_iter=60)
    ic = ctx.get_state()
    ctx.close()
    assert np.all(iters >= 1) and np.all(np.isfinite(ic))
    tape = tape_for(cols, len(cols))
    out = run_job(cols, ic, tape=tape, out_stride=10)
    again = run_job(cols, ic, tape=tape, out_stride=10)
    assert same_bits(out, again) is None
    alone = run_job(cols, ic, tape=tape, out_stride=10, sel=[5, 23])
    assert same_bits(out, alone, sel=[5, 23]) is None
    assert np.all(np.isfinite(out["state"]))
    w = out["wscal"]
    assert np.all((w[:, :, pack.W_WTD_EST] >= 0) & (w[:, :, pack.W_WTD_EST] <= nx - 1))
    assert np.all(out["counters"][:, pack.C_NWINDOWS] == 39)
    oracle_sample(cols, tape, ic, out, sample=[0, 11], rows=12, cells=2, frac=0.9)
