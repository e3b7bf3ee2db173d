"""
GPU tests added in round 2 (run on the B200 box, all through the C-ABI):

* BASELINE.json configs[2] at its full shape (65,536 columns, 3,650-row forcing,
  Philox noise): slicing / sharding / schedule independence, invariants, and a
  sample of members against the oracle fed with the SAME Philox fields;
* configs[4] at 4,096 columns x nx = 1001;
* integer / bit-exact pieces on the device: find_wtd on the reference's four test
  vectors, the Philox stream (moments, KS, stream disjointness);
* failed attempts under time-sliced scheduling (the damped base noise field is
  handed from SM to SM) are schedule independent;
* the one-pass Jacobian differences (jac_fused) are bit-identical to the evaluation per group;
* a noise tape that is too short fails loudly;
* per-column forcing series, ensemble statistics incl. theta moments, the
  multi-GPU entry `EnsembleSimulation.from_global`, the command-line entry point.
"""
import json
import os

import numpy as np
import pytest

import oracle
from hydromodel_b200 import capi, pack
from hydromodel_b200.ensemble import EnsembleSimulation
from hydromodel_b200.synthetic import (default_params, ensemble_columns, ensemble_member,
                                       fine_grid_site, synthetic_forcing, write_forcing_csv)
from helpers import SITE, default_column, tol_scaled

pytestmark = pytest.mark.gpu

IC = os.path.join(os.path.dirname(SITE), "ic_well10_default.npy")
SEED = 20240607


def open_ctx(ens, n_cols=None, first=0, device=0):
    n = ens["par"].shape[0] if n_cols is None else n_cols
    nx = ens["x"].size
    ctx = capi.Context(nx, n, device)
    grp, ng = pack.fd_groups(nx)
    ctx.set_grid(ens["x"], grp, ng)
    ctx.set_columns(ens["par"][first:first + n], ens["stat"][first:first + n])
    f = ens["forcing"]
    ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])
    return ctx, grp


# --------------------------------------------------------------------------- configs[2]
def test_config2_full_shape_slicing_sharding_and_oracle_sample(golden):
    """65,536 members (n / a0 / psi_sat / Theta_Max / precipitation multiplier), 3,650-row table,
    Philox noise -- the first 60 windows.  (a) the job integrated in two row slices on one context
    equals, bit for bit, a 4,096-column shard of it (global ids 20,000 ..) integrated in one launch
    on another context with the plain schedule; (b) invariants; (c) eight members against the
    oracle driven by the same Philox fields, free-running bars from noise_floor.npz."""
    C, T, SH0, SHN = 65536, 61, 20000, 4096
    table = synthetic_forcing(3650)
    base, cols = ensemble_columns(default_params(SITE), 0, C, perturb_theta=True)
    ens = pack.pack_ensemble(base, table, cols)
    assert ens["n_slow"] == 0 and ens["par"].shape == (C, pack.NPARAM)
    ic = np.load(IC)
    nx = ens["x"].size

    ctx, grp = open_ctx(ens)
    ctx.set_noise_philox(SEED, 0)
    ctx.set_output(60)                       # output rows 0 and 60 fall into the 60 windows
    ctx.set_state(np.tile(ic, (C, 1)))
    ctx.reset_counters()
    ctx.run(0, 31)
    mid = ctx.get_state()
    ctx.set_state(mid)                       # the host round trip of a sliced run
    ctx.run(31, T)
    state = ctx.get_state()
    wscal = ctx.get_wscal(rows=(0, T))
    fields = ctx.get_fields(out_rows=(0, 2))
    counters = ctx.get_counters()
    ctx.close()

    # (b) invariants
    assert np.all(np.isfinite(state)) and np.all(np.isfinite(wscal))
    wtd = wscal[:, :, pack.W_WTD_EST]
    assert np.all((wtd >= 0) & (wtd <= nx - 1))
    assert np.all(counters[:, pack.C_NWINDOWS] == T - 1)
    assert np.mean(wscal[:, 1:, pack.W_STATUS] == 0) >= 0.9999
    th, K, kb = fields[:, :, pack.F_THETA], fields[:, :, pack.F_K], fields[:, :, pack.F_KBKG]
    por, thr = ens["stat"][:, None, pack.S_PHI_NODE, :], ens["par"][:, None, None, pack.P_THETA_RES]
    assert np.all(th <= por * (1 + 1e-12)) and np.all(th >= thr * (1 - 1e-12))
    assert np.all(K <= kb * (1 + 1e-12)) and np.all(K >= 0.0)
    obs = ens["forcing"]["wtd_idx"][:T]
    assert np.mean(np.abs(wtd[:, 1:] - obs[None, 1:]) <= 25) >= 0.99

    # (a) a shard on its own context, one launch, plain chunk-major schedule
    sh, _ = open_ctx(ens, SHN, SH0)
    sh.set_schedule(False)
    sh.set_noise_philox(SEED, SH0)
    sh.set_output(60)
    sh.set_state(np.tile(ic, (SHN, 1)))
    sh.reset_counters()
    sh.run(0, T)
    assert np.array_equal(sh.get_state(), state[SH0:SH0 + SHN])
    assert np.array_equal(sh.get_wscal(rows=(0, T)), wscal[SH0:SH0 + SHN])
    assert np.array_equal(sh.get_fields(out_rows=(0, 2)), fields[SH0:SH0 + SHN])
    assert np.array_equal(sh.get_counters(), counters[SH0:SH0 + SHN])

    # (c) oracle sample: same parameters, same Philox fields as a host tape
    nf = golden("noise_floor.npz")
    psi_bar = 2.0 * max(nf["free_e12_units"][:T].max(), nf["free_e14_units"][:T].max())
    f = ens["forcing"]
    rows = np.arange(T)
    worst = 0.0
    for m in (0, 1, 2, 3, 4093, 4094, 4095, 1717):
        g = SH0 + m
        mult = ens["par"][g, pack.P_PRECIP_MULT]
        refresh = ((mult * f["precipitation_cm"][:T] > 0.5) | (rows % 48 == 0)) & (rows > 0) & (f["wtd_idx"][:T] >= 0)
        tape = np.stack([sh.philox_normals(SEED, g, 0)] + [sh.philox_normals(SEED, g, int(r)) for r in rows[refresh]])
        o_psi, o_wscal, _, o_stats, _ = oracle.run(ens["x"], grp, ens["par"][g], ens["stat"][g],
                                                   f["precipitation_cm"][:T], f["wtd_idx"][:T], f["daylight"][:T],
                                                   f["atm"][:T], tape, ic, out_stride=30)
        err = tol_scaled(state[g], o_psi)
        worst = max(worst, err)
        assert err < psi_bar, (g, err)
        d = np.abs(wscal[g, :, pack.W_WTD_EST] - o_wscal[:, pack.W_WTD_EST])
        assert d.max() <= 1 and np.mean(d == 0) >= 0.95, g
        assert abs(counters[g, pack.C_NFEV] - o_stats[pack.C_NFEV]) <= 0.05 * o_stats[pack.C_NFEV]
    print("configs[2] oracle sample: worst psi difference %.3f tolerance units (bar %.2f)" % (worst, psi_bar))
    sh.close()


# --------------------------------------------------------------------------- configs[4]
def test_config4_fine_grid_4096_columns(tmp_path):
    """4,096 columns on a 1,001-node grid under the stiff heavy-rain forcing (12 rows): invariants,
    and a 64-column shard on its own context gives the same bits."""
    site = fine_grid_site(str(tmp_path / "site.json"))
    table = synthetic_forcing(365, stiff=True).iloc[:13].reset_index(drop=True)
    C, T = 4096, 13
    base, cols = ensemble_columns(default_params(site), 0, C)
    ens = pack.pack_ensemble(base, table, cols)
    nx = ens["x"].size
    assert nx == 1001
    z = ens["x"]
    ic = np.where(z > 800.0, z - 800.0, -0.35 * (800.0 - z) - 1.0)
    ctx, _ = open_ctx(ens)
    ctx.set_noise_philox(SEED, 0)
    ctx.set_output(6)
    ctx.set_state(np.tile(ic, (C, 1)))
    ctx.reset_counters()
    ctx.run(0, T)
    state, wscal, counters = ctx.get_state(), ctx.get_wscal(), ctx.get_counters()
    fields = ctx.get_fields()
    ctx.close()
    assert np.all(np.isfinite(state)) and np.all(counters[:, pack.C_NWINDOWS] == T - 1)
    assert np.mean(wscal[:, 1:, pack.W_STATUS] == 0) >= 0.999
    th = fields[:, :, pack.F_THETA]
    assert np.all(th <= ens["stat"][:, None, pack.S_PHI_NODE, :] * (1 + 1e-12))
    assert np.all(th >= ens["par"][:, None, None, pack.P_THETA_RES] * (1 - 1e-12))
    sh, _ = open_ctx(ens, 64, 1000)
    sh.set_noise_philox(SEED, 1000)
    sh.set_output(6)
    sh.set_state(np.tile(ic, (64, 1)))
    sh.run(0, T)
    assert np.array_equal(sh.get_state(), state[1000:1064])
    assert np.array_equal(sh.get_wscal(), wscal[1000:1064])
    sh.close()


# --------------------------------------------------------------------------- integer pieces
REF_FIND_WTD = (        # code/tests/test_utilities.py:56-86, exact
    ([0, 0, 0, 0, 0, 0, 0, 0, 0, 0], 9),
    ([1, 1, 1, 1, 1, 1, 1, 1, 1, 1], 0),
    ([0, 0, 0, 0, 0, 0, 1, 1, 1, 1], 6),
    ([0, 1, 0, 1, 0, 0, 0, 1, 1, 1], 7),
)


def test_find_wtd_on_device_reference_vectors():
    """find_wtd (utilities.py:56-99) as the kernels evaluate it (block reduction over
    `psi >= psi_sat`): the reference's four vectors, and random masks on 401 / 1001 nodes
    against the host function."""
    from hydromodel_b200.src.utilities import find_wtd
    rng = np.random.default_rng(8)
    for nx in (10, 401, 1001):
        ctx = capi.Context(nx, 1)
        grp, ng = pack.fd_groups(nx)
        ctx.set_grid(np.arange(nx) * 5.0, grp, ng)
        par = np.zeros((1, pack.NPARAM))
        par[0, pack.P_PSI_SAT] = -0.5
        ctx.set_columns(par, np.zeros((1, pack.NSTATIC, nx)))
        masks = [np.array(v, dtype=bool) for v, _ in REF_FIND_WTD] if nx == 10 else \
            [rng.random(nx) < p for p in (0.0, 1.0, 0.5, 0.02, 0.98)] + \
            [np.arange(nx) >= k for k in (0, 1, nx // 2, nx - 1, nx)]
        for k, sat in enumerate(masks):
            psi = np.where(sat, 3.0, -7.0)
            got = ctx.find_wtd(0, psi)
            assert got == find_wtd(sat) == oracle.find_wtd(sat)
            if nx == 10:
                assert got == REF_FIND_WTD[k][1]
        ctx.close()


def test_philox_stream_statistics_and_disjointness():
    """Philox4x32-10 + Box-Muller as the kernels call it: 1e6 draws of one stream are standard
    normal (mean, variance, skewness, kurtosis, Kolmogorov-Smirnov), and the streams of different
    (column, window) keys -- incl. the spin-up key -1 and the base-field key 0 -- share no values
    and are uncorrelated."""
    from scipy import stats
    ctx = capi.Context(16, 1)
    n = 1_000_000
    x = ctx.philox_normals(SEED, 7, 3, n)
    assert abs(x.mean()) < 4.0 / np.sqrt(n)
    assert abs(x.var() - 1.0) < 4.0 * np.sqrt(2.0 / n)
    assert abs(stats.skew(x)) < 4.0 * np.sqrt(6.0 / n) and abs(stats.kurtosis(x)) < 4.0 * np.sqrt(24.0 / n)
    assert stats.kstest(x, "norm").pvalue > 1e-3
    assert np.unique(x).size > n - 10
    m = 4096
    keys = [(0, -1), (0, 0), (0, 1), (0, 2), (1, -1), (1, 0), (1, 1), (2 ** 32, 0), (2 ** 32 + 1, 1), (65535, 3649)]
    draws = np.stack([ctx.philox_normals(SEED, c, w, m) for c, w in keys])
    assert np.array_equal(draws[2], ctx.philox_normals(SEED, 0, 1, m))          # reproducible
    assert not np.array_equal(draws[2], ctx.philox_normals(SEED + 1, 0, 1, m))  # keyed by the seed
    for a in range(len(keys)):
        for b in range(a + 1, len(keys)):
            assert np.intersect1d(draws[a], draws[b]).size == 0, (keys[a], keys[b])
            assert abs(np.corrcoef(draws[a], draws[b])[0, 1]) < 5.0 / np.sqrt(m), (keys[a], keys[b])
    # the (cos, sin) pair of one counter is not reused across neighbouring indices of a stream
    assert abs(np.corrcoef(draws[2][0::2], draws[2][1::2])[0, 1]) < 5.0 / np.sqrt(m / 2)
    ctx.close()


# --------------------------------------------------------------------------- failures x scheduling
def _run_damped(chunk, adaptive, ens, tape, T):
    """A base noise field of amplitude 1e6 makes K_bkg overflow in the unsaturated zone: every
    attempt of the first windows fails fast (non-finite RHS -> step-size underflow), the field is
    damped by 0.8 per failed attempt IN PLACE (richards_pde.py:519-523) and handed from window to
    window until it is small enough for the solver to succeed."""
    old = os.environ.pop("HB_CHUNK", None)
    if chunk:
        os.environ["HB_CHUNK"] = str(chunk)
    try:
        n_cols = ens["par"].shape[0]
        ctx, _ = open_ctx(ens)
        ctx.set_schedule(adaptive)
        ic = np.tile(np.load(IC), (n_cols, 1))
        outs = []
        for _ in range(2 if adaptive else 1):      # second launch: the plan knows the expensive columns
            ctx.set_noise_tape(tape)
            ctx.set_output(8)
            ctx.set_state(ic)
            ctx.reset_counters()
            ctx.run(0, T)
            outs.append((ctx.get_state(), ctx.get_wscal(), ctx.get_fields(), ctx.get_counters()))
        ctx.close()
        for a, b in zip(outs[0], outs[-1]):
            assert np.array_equal(a, b, equal_nan=True)
        return outs[0]
    finally:
        os.environ.pop("HB_CHUNK", None)
        if old is not None:
            os.environ["HB_CHUNK"] = old


def test_failed_attempts_are_schedule_independent():
    """Failed attempts under time-sliced scheduling.  With one-row chunks every row of a column
    may run on another SM, so the damped base field, the state and the tape position travel
    through global memory: the results must equal, bit for bit, the run with 16-row chunks and the
    run under the cost-aware plan (more columns than resident CTAs, so chunks do migrate).  Two
    columns are also compared with the oracle: same failed windows, same number of failed attempts
    (+-5), same water-table node at the end."""
    n_cols, T = 320, 25
    table = synthetic_forcing(365).iloc[:T].reset_index(drop=True)
    base, cols = ensemble_columns(default_params(SITE), 0, n_cols)
    ens = pack.pack_ensemble(base, table, cols)
    f = ens["forcing"]
    rows = np.arange(T)
    slots = 1 + int((((f["precipitation_cm"] > 0.5) | (rows % 48 == 0)) & (rows > 0)).sum())
    rng = np.random.default_rng(77)
    tape = rng.standard_normal((n_cols, slots, 401))
    tape[:, 0] *= 1.0e6
    ref = _run_damped(None, False, ens, tape, T)
    fails, status = ref[3][:, pack.C_NFAIL], ref[1][:, :, pack.W_STATUS]
    print("failed attempts: %d in %d columns (max %d per column), failed windows %d, RHS evaluations per column "
          "median %.0f" % (fails.sum(), (fails > 0).sum(), fails.max(), (status == 1).sum(),
                           np.median(ref[3][:, pack.C_NRHS])))
    assert (fails >= 10).all() and (status[:, -5:] == 0).all(), "every column must fail first and recover"
    for chunk, adaptive in ((1, False), (3, True)):
        got = _run_damped(chunk, adaptive, ens, tape, T)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b, equal_nan=True), (chunk, adaptive)
    assert np.all(np.isfinite(ref[0])), "a failed attempt returns the last accepted state, never NaN"
    grp, _ = pack.fd_groups(401)
    ic = np.load(IC)
    for c in (0, 171):
        o_psi, o_wscal, _, o_stats, _ = oracle.run(ens["x"], grp, ens["par"][c], ens["stat"][c], f["precipitation_cm"],
                                                   f["wtd_idx"], f["daylight"], f["atm"], tape[c], ic, out_stride=8)
        assert np.mean(o_wscal[:, pack.W_STATUS] == status[c]) >= 0.9, c
        assert abs(int(o_stats[pack.C_NFAIL]) - int(fails[c])) <= 5, c
        assert abs(o_wscal[-1, pack.W_WTD_EST] - ref[1][c, -1, pack.W_WTD_EST]) <= 1, c


def _both_jacobian_paths(work):
    """Run `work()` with the grouped Jacobian differences made in one pass (the default,
    jac_fused) and with one right-hand-side evaluation per colour group (HB_JAC_SEQ=1, scipy's own
    sequence of evaluations); returns the two result tuples."""
    res = []
    for seq in (False, True):
        os.environ.pop("HB_JAC_SEQ", None)
        if seq:
            os.environ["HB_JAC_SEQ"] = "1"
        try:
            res.append(work())
        finally:
            os.environ.pop("HB_JAC_SEQ", None)
    return res


def test_fused_jacobian_is_bit_identical(tmp_path):
    """num_jac's first pass over the colour groups (SCIPY common.py:385-443) is made in ONE pass on
    the device: two closure evaluations per midpoint (own node perturbed / right neighbour
    perturbed) instead of one whole right-hand-side evaluation per group, the three rows of every
    column assembled from them.  Every number must be the one the evaluation per group produces:
    states, per-row scalars, output fields and all solver counters are compared bit for bit over
    (1) a mixed ensemble (both closures, all porosity profiles, lateral-flow sink on) with daily
    rows, (2) the same without the lateral-flow sink, (3) half-hourly rows with hydraulic lift at
    night (fused) and the root-uptake sink by day (evaluation per group), (4) failed attempts with
    the damped noise field, (5) the spin-up kernel, (6) the default column (n = 2 closure),
    (7) the kernel variants of larger grids (701 and 1,001 nodes)."""
    n_cols, T = 192, 31
    table = synthetic_forcing(365).iloc[:T].reset_index(drop=True)
    base, cols = ensemble_columns(default_params(SITE), 4096, n_cols, mixed=True, perturb_theta=True)
    ens = pack.pack_ensemble(base, table, cols)
    ic = np.tile(np.load(IC), (n_cols, 1))

    def run(e, rows, tape=None, stride=10):
        ctx, _ = open_ctx(e)
        if tape is None:
            ctx.set_noise_philox(SEED, 4096)
        else:
            ctx.set_noise_tape(tape)
        ctx.set_output(stride)
        ctx.set_state(ic[:e["par"].shape[0]])
        ctx.reset_counters()
        ctx.run(0, rows)
        out = (ctx.get_state(), ctx.get_wscal(), ctx.get_fields(), ctx.get_counters())
        ctx.close()
        return out

    def same(a, b, what):
        for u, v, name in zip(a, b, ("state", "per-row scalars", "fields", "counters")):
            assert np.array_equal(u, v, equal_nan=True), (what, name)

    # (1) mixed ensemble, default flags (LF on; ET flag set but daily rows are night rows)
    a, b = _both_jacobian_paths(lambda: run(ens, T))
    same(a, b, "mixed ensemble")
    njev = a[3][:, pack.C_NJEV].sum() / a[3][:, pack.C_NWINDOWS].sum()
    assert njev > 1.2, "the workload must build Jacobians"
    # (2) no lateral-flow sink
    e2 = dict(ens); e2["par"] = ens["par"].copy(); e2["par"][:, pack.P_FLAGS] = 0.0
    a2, b2 = _both_jacobian_paths(lambda: run(e2, T))
    same(a2, b2, "no sinks")
    assert not np.array_equal(a[0], a2[0])
    # (3) half-hourly rows, hydraulic lift + root uptake (gentle demand keeps the daylight rows cheap)
    hh = synthetic_forcing(40, half_hourly=True)
    e3 = pack.pack_ensemble(base, hh, {k: v[:48] for k, v in cols.items()})
    e3["par"] = e3["par"].copy(); e3["par"][:, pack.P_FLAGS] = 7.0          # LF | ET | HLIFT
    e3["forcing"] = dict(e3["forcing"]); e3["forcing"]["atm"] = e3["forcing"]["atm"] * 1.0e-3
    assert e3["forcing"]["daylight"].any() and not e3["forcing"]["daylight"].all()
    a3, b3 = _both_jacobian_paths(lambda: run(e3, 40, stride=8))
    same(a3, b3, "half-hourly, HLIFT + ET")
    # (4) failed attempts (noise field damped in place, attempts restarted)
    e4 = {k: (v[:64] if k in ("par", "stat") else v) for k, v in ens.items()}
    f = ens["forcing"]
    rows = np.arange(T)
    slots = 1 + int((((f["precipitation_cm"] > 0.5) | (rows % 48 == 0)) & (rows > 0)).sum())
    tape = np.random.default_rng(78).standard_normal((64, slots, 401))
    tape[:, 0] *= 1.0e6
    a4, b4 = _both_jacobian_paths(lambda: run(e4, 20, tape=tape))
    same(a4, b4, "failed attempts")
    assert (a4[3][:, pack.C_NFAIL] >= 5).sum() >= 16       # (van Genuchten members ignore the noise field)
    # (5) spin-up kernel, (6) default column (n == 2 closure) in the window loop
    col = default_column()

    def spin():
        ctx = capi.Context(401, 8)
        grp, ng = pack.fd_groups(401)
        ctx.set_grid(col["x"], grp, ng)
        ctx.set_columns(np.tile(col["par"], (8, 1)), np.tile(col["stat"], (8, 1, 1)))
        fc = col["forcing"]
        ctx.set_forcing(fc["precipitation_cm"][:25], fc["wtd_idx"][:25], fc["daylight"][:25], fc["atm"][:25])
        ctx.set_noise_philox(SEED, 0)
        psi0 = np.tile(np.load(IC), (8, 1)) * np.linspace(0.9, 1.1, 8)[:, None]
        ctx.set_state(psi0)
        iters = ctx.spinup(fc["zWtd_cm"][0], None, max_iter=25)
        spun = ctx.get_state()
        ctx.set_output(6)
        ctx.reset_counters()
        ctx.run(0, 25)
        out = (spun, ctx.get_wscal(), ctx.get_fields(), ctx.get_counters(), iters, ctx.get_state())
        ctx.close()
        return out

    a5, b5 = _both_jacobian_paths(spin)
    for u, v in zip(a5, b5):
        assert np.array_equal(u, v, equal_nan=True)
    assert col["par"][pack.P_N] == 2.0
    # (7) the other kernel variants: 701 nodes (512 threads x 2 cells, run-time vector length, generic
    # tridiagonal solver) and 1,001 nodes (512 x 3) under the stiff forcing
    for depth, want_nx in ((3500.0, 701), (5000.0, 1001)):
        site = fine_grid_site(str(tmp_path / ("site%d.json" % want_nx)), max_depth=depth)
        tb = synthetic_forcing(365, stiff=True).iloc[:9].reset_index(drop=True)
        bs, cs = ensemble_columns(default_params(site), 0, 24)
        eg = pack.pack_ensemble(bs, tb, cs)
        z = eg["x"]
        assert z.size == want_nx
        icg = np.tile(np.where(z > 800.0, z - 800.0, -0.35 * (800.0 - z) - 1.0), (24, 1))

        def fine():
            ctx, _ = open_ctx(eg)
            ctx.set_noise_philox(SEED, 0)
            ctx.set_output(4)
            ctx.set_state(icg)
            ctx.reset_counters()
            ctx.run(0, 9)
            out = (ctx.get_state(), ctx.get_wscal(), ctx.get_fields(), ctx.get_counters())
            ctx.close()
            return out

        a7, b7 = _both_jacobian_paths(fine)
        same(a7, b7, "nx = %d" % want_nx)
        assert a7[3][:, pack.C_NJEV].sum() >= 24 * 8


def test_short_noise_tape_fails_loudly():
    """A column that refreshes its field more often than the tape has slots must not silently
    reuse the last slot: the next getter reports it."""
    col = default_column()
    f = col["forcing"]
    T = 40
    ctx = capi.Context(col["x"].size, 2)
    grp, ng = pack.fd_groups(col["x"].size)
    ctx.set_grid(col["x"], grp, ng)
    par = np.tile(col["par"], (2, 1))
    par[1, pack.P_PRECIP_MULT] = 3.0           # refreshes on more rows than column 0
    ctx.set_columns(par, np.tile(col["stat"], (2, 1, 1)))
    ctx.set_forcing(f["precipitation_cm"][:T], f["wtd_idx"][:T], f["daylight"][:T], f["atm"][:T])
    rows = np.arange(T)
    need = [1 + int((((m * f["precipitation_cm"][:T] > 0.5) | (rows % 48 == 0)) & (rows > 0)).sum()) for m in (1.0, 3.0)]
    assert need[1] > need[0]
    rng = np.random.default_rng(0)
    ctx.set_noise_tape(rng.standard_normal((2, need[0], col["x"].size)))
    ctx.set_state(np.tile(np.load(IC), (2, 1)))
    ctx.run(0, T)
    with pytest.raises(capi.HydroB200Error, match="noise tape too short"):
        ctx.get_state()
    ctx.set_noise_tape(rng.standard_normal((2, need[1], col["x"].size)))
    ctx.set_state(np.tile(np.load(IC), (2, 1)))
    ctx.run(0, T)
    ctx.get_state()
    ctx.close()


# --------------------------------------------------------------------------- ensemble API
def test_per_column_forcing_series_and_statistics():
    """Two forcing tables (the second with 1.5x the rain) mapped to the columns by
    `series_of_col`: every column equals the same column of a single-table ensemble; the device
    statistics (count, WTD moments, theta moments) equal NumPy on the outputs."""
    T = 33
    t0 = synthetic_forcing(365).iloc[:T].reset_index(drop=True)
    t1 = t0.copy()
    t1["Precipitation_cm"] = np.round(t1["Precipitation_cm"] * 1.5, 2)
    base, cols = ensemble_columns(default_params(SITE), 0, 12, perturb_theta=True)
    ser = np.array([0, 1] * 6)
    ic = np.load(IC)
    outs = {}
    for tag, data, s in (("both", [t0, t1], ser), ("a", t0, None), ("b", t1, None)):
        ens = EnsembleSimulation(tag, seed=5, noise="philox")
        ens.setupEnsemble(base, data, cols, series_of_col=s)
        ens.set_initial_conditions(ic)
        out = ens.run(out_stride=8)
        outs[tag] = {k: np.array(v) for k, v in out.items()}
        if tag == "both":
            st = ens.statistics()
            wtd_node = out["wtd_est_cm"] / 5.0
            assert np.all(st["count"] == 12) and st["n_columns"] == 12
            assert np.allclose(st["wtd_node_mean"], wtd_node.mean(axis=0), rtol=1e-13)
            assert np.allclose(st["wtd_node_std"], wtd_node.std(axis=0), rtol=1e-6, atol=1e-6)
            assert np.allclose(st["theta_mean"], out["theta_vol"].mean(axis=0), rtol=1e-13)
            # (std from sums of squares: cancellation where all columns share a value)
            assert np.allclose(st["theta_std"], out["theta_vol"].std(axis=0), rtol=1e-4, atol=1e-7)
        ens.close()
    for key in ("wtd_est_cm", "theta_vol", "psi_press", "lateral_flow"):
        assert np.array_equal(outs["both"][key][0::2], outs["a"][key][0::2]), key
        assert np.array_equal(outs["both"][key][1::2], outs["b"][key][1::2]), key
    assert not np.array_equal(outs["a"]["psi_press"][1], outs["b"]["psi_press"][1])


def test_from_global_shares_tile_the_ensemble():
    """`EnsembleSimulation.from_global` for world = 3 on one GPU: the three ranks' results tile
    the result of the undivided ensemble (global member ids key members and noise)."""
    T, C = 25, 50
    table = synthetic_forcing(365).iloc[:T].reset_index(drop=True)
    base = default_params(SITE)
    ic = np.load(IC)

    def members(first, count):
        return ensemble_columns(base, first, count, perturb_theta=True)

    whole = EnsembleSimulation.from_global(None, table, members, C, rank=0, world=1, seed=9, device=0)
    whole.set_initial_conditions(ic)
    ref = {k: np.array(v) for k, v in whole.run(out_stride=12).items()}
    whole.close()
    parts = []
    for r in range(3):
        sim = EnsembleSimulation.from_global(None, table, members, C, rank=r, world=3, seed=9, device=0)
        assert sim.n_global == C
        sim.set_initial_conditions(ic)
        parts.append({k: np.array(v) for k, v in sim.run(out_stride=12).items()})
        sim.close()
    for key in ("wtd_est_cm", "theta_vol", "psi_press", "status"):
        assert np.array_equal(np.concatenate([p[key] for p in parts]), ref[key]), key


def test_command_line_entry_point(tmp_path, capsys):
    """berkeley_hydro_main.main (code/berkeley_hydro_main.py:65-146): JSON + CSV in, <Output_Name>.h5
    out with the reference's dataset names and shapes; every failure ends with exit code 1."""
    from hydromodel_b200 import berkeley_hydro_main as cli
    from hydromodel_b200.src.simulation import loadResults
    T = 21
    csv = tmp_path / "water.csv"
    write_forcing_csv(synthetic_forcing(365).iloc[:T], str(csv))
    ic_csv = tmp_path / "ic.csv"
    np.savetxt(str(ic_csv), np.load(IC), fmt="%.17g")
    params = default_params(SITE, IC_Filename=str(ic_csv), Data_Filename=str(csv),
                            Output_Name=str(tmp_path / "Sim_cli"))
    pfile = tmp_path / "params.json"
    pfile.write_text(json.dumps(params))
    cli.main(str(pfile), None)                       # CSV taken from Data_Filename
    res = loadResults(str(tmp_path / "Sim_cli.h5"))
    for key in ("K_hrc", "K_bkg", "S_eff", "psi_press", "theta_vol"):
        assert res[key].shape == (T, 401), key
    assert res["abs_error"].shape == (T,) and res["wtd_est_cm"].shape == (T,)
    assert res["lateral_flow"].shape == (T - 1,) and res["transpiration"].shape == (T - 1,)
    assert np.all(np.isfinite(res["psi_press"]))
    # exit-code-1 contract
    for bad in ("no_params", "missing_key", "missing_csv"):
        args = {"no_params": (None, None),
                "missing_key": (str(tmp_path / "bad.json"), str(csv)),
                "missing_csv": (str(pfile), str(tmp_path / "nope.csv"))}[bad]
        if bad == "missing_key":
            broken = dict(params)
            del broken["Soil_Properties"]
            (tmp_path / "bad.json").write_text(json.dumps(broken))
        with pytest.raises(SystemExit) as exc:
            cli.main(*args)
        assert exc.value.code == 1, bad


# --------------------------------------------------------------------------- ET windows
def test_halfhourly_daylight_windows_against_reference_tape(golden):
    """Reference-taped half-hourly run with the default flags (ET on), gentle demand: 1 night +
    23 daylight windows with the root-water-uptake sink active (block prefix sum + six block
    reductions per RHS evaluation), teacher-forced against the REFERENCE's tape: end state,
    counters, transpiration and lateral flow of the last RHS evaluation (arg_out)."""
    from helpers import column_from_json
    g = golden("windows_halfhourly_gentle.npz")
    col = column_from_json(g["params_json"])
    nx = col["x"].size
    ctx = capi.Context(nx, 1)
    grp, ng = pack.fd_groups(nx)
    ctx.set_grid(col["x"], grp, ng)
    ctx.set_columns(col["par"][None], col["stat"][None])
    errs, same, n_day = [], [], 0
    for w in range(g["y0"].shape[0]):
        nr = g["n_rnd"][w].copy()
        day = int(6 <= g["hour"][w] < 18)
        t0 = float(g["t0"][w])
        y, st, status, lat = ctx.solve_window(0, t0, t0 + 1.0, g["y0"][w], nr, float(g["precip"][w]),
                                              int(g["wtd"][w]), day, 0, atm=float(g["atm"][w]))
        assert status == 0
        errs.append(tol_scaled(y, g["y_end"][w]))
        same.append(st["nfev"] == g["nfev"][w] and st["njev"] == g["njev"][w] and st["nlu"] == g["nlu"][w]
                    and st["nsteps"] == g["nsteps"][w])
        assert ctx.last_transpiration == pytest.approx(float(g["transpiration"][w]), rel=1e-9, abs=1e-300), w
        assert lat == pytest.approx(float(g["lateral_flow"][w]), rel=1e-6, abs=1e-9), w
        n_day += day
    ctx.close()
    errs, same = np.array(errs), np.array(same)
    print("gpu half-hourly windows vs reference tape: %d daylight; counters-match %.3f; err median %.2e max %.2e"
          % (n_day, same.mean(), np.median(errs), errs.max()))
    assert n_day >= 20
    assert same.mean() >= 0.85
    assert np.median(errs) < 1e-3 and errs.max() < 5.0


def test_halfhourly_default_demand_windows_behave_like_the_reference_tape(golden):
    """The same run with the DEFAULT atmospheric demand (tests/golden/windows_halfhourly.npz: what
    the reference got through in 90 minutes of wall clock -- 10 solve_ivp calls): the root-uptake
    sink is discontinuous in the state, and the reference itself needs 548,722 RHS evaluations for
    the first daylight window, fails two attempts of the next one and all five of the third.
    Trajectories in this regime are chaotic (no two implementations, nor the reference twice, take
    the same 95,651 steps), so what is pinned is what is deterministic: the night window to the
    usual bar, and for every daylight window started from the reference's state a finite end
    state, the tape's transpiration when the window is solved, and the reference's failure
    semantics (noise damped by 0.8 per failed attempt); the cost on both sides is printed."""
    from helpers import column_from_json
    g = golden("windows_halfhourly.npz")
    col = column_from_json(g["params_json"])
    nx = col["x"].size
    ctx = capi.Context(nx, 1)
    grp, ng = pack.fd_groups(nx)
    ctx.set_grid(col["x"], grp, ng)
    ctx.set_columns(col["par"][None], col["stat"][None])
    t0s = g["t0"]
    first = [k for k in range(t0s.size) if k == 0 or t0s[k] != t0s[k - 1]]       # first attempt of each window
    report = []
    for k in first:
        nr = g["n_rnd"][k].copy()
        day = int(6 <= g["hour"][k] < 18)
        y, st, status, lat = ctx.solve_window(0, float(t0s[k]), float(t0s[k]) + 1.0, g["y0"][k], nr, float(g["precip"][k]),
                                              int(g["wtd"][k]), day, 0, atm=float(g["atm"][k]))
        ref_nfev = int(g["nfev"][t0s == t0s[k]].sum())
        report.append((int(g["hour"][k]), day, status, st["nfail"], st["nfev"], ref_nfev))
        assert np.all(np.isfinite(y)), k
        assert np.allclose(nr, g["n_rnd"][k] * 0.8 ** st["nfail"], rtol=1e-14)      # quirk Q4
        if not day:
            assert status == 0 and tol_scaled(y, g["y_end"][k]) < 0.1
            assert ctx.last_transpiration == 0.0
        else:
            if status == 0:
                assert ctx.last_transpiration == pytest.approx(float(g["transpiration"][k]), rel=1e-9)
    ctx.close()
    print("default-demand half-hourly windows (hour, daylight, status, failed attempts, nfev device, nfev reference): %s"
          % report)
    assert sum(r[1] for r in report) >= 3


# --------------------------------------------------------------------------- fast mode
def test_fast_mode_against_parity_mode():
    """SURVEY 8(f).4, explicit flag, never the default: windows whose forcing row and noise field
    repeat the previous window's are integrated as the continuation of that window (order, step
    size, difference table and Jacobian carried on) and the finite-difference Jacobian uses 3
    colours without retry pass.  It takes other steps than scipy's restarted controller, so it is
    judged like a second integrator at the same rtol = atol = 1e-3 -- over 120 windows of the
    configs[1] workload (256 members, shared IC, Philox noise) against the parity mode:
    water-table node identical in >= 97 % of the rows and never more than 1 cell off; theta within
    3e-3 relative at every output row and node (the parity mode's own reproducibility over such a
    run is 5e-4, noise_floor.npz); psi, which moves by orders of magnitude across a wetting front,
    within 30 solver-tolerance units in the median column -- with fewer RHS evaluations."""
    C, T = 256, 121
    table = synthetic_forcing(365).iloc[:T].reset_index(drop=True)
    base, cols = ensemble_columns(default_params(SITE), 0, C)
    ens = pack.pack_ensemble(base, table, cols)
    ic = np.load(IC)
    res = {}
    for mode in ("parity", "fast", "fast_f32"):
        ctx, _ = open_ctx(ens)
        if mode != "parity":
            ctx.set_fast_mode(cont=True, jac3=True, f32=(mode == "fast_f32"))
        ctx.set_noise_philox(SEED, 0)
        ctx.set_output(20)
        ctx.set_state(np.tile(ic, (C, 1)))
        ctx.reset_counters()
        ctx.run(0, T)
        res[mode] = (ctx.get_state(), ctx.get_wscal(), ctx.get_fields(), ctx.get_counters())
        ctx.close()
    p_psi, p_ws, p_f, p_cn = res["parity"]
    for mode in ("fast", "fast_f32"):
        # HB_FAST_F32 on top: the closure of the Newton evaluations in single precision (errors of
        # ~1e-6 in theta, K, C against a Newton tolerance of 0.03 solver units) -- same bars
        f_psi, f_ws, f_f, f_cn = res[mode]
        d = np.abs(p_ws[:, :, pack.W_WTD_EST] - f_ws[:, :, pack.W_WTD_EST])
        psi_units = np.max(np.abs(f_psi - p_psi) / (1e-3 + 1e-3 * np.abs(p_psi)), axis=1)
        th_rel = np.max(np.abs(f_f[:, :, pack.F_THETA] - p_f[:, :, pack.F_THETA]) / p_f[:, :, pack.F_THETA])
        rhs_ratio = f_cn[:, pack.C_NRHS].sum() / p_cn[:, pack.C_NRHS].sum()
        print("%s vs parity: wtd node identical %.4f (max diff %d), psi units median %.3f max %.3f, theta rel max "
              "%.2e, RHS evaluations x%.3f, steps x%.3f, Jacobians x%.3f, failed attempts %d vs %d" %
              (mode, np.mean(d == 0), d.max(), np.median(psi_units), psi_units.max(), th_rel, rhs_ratio,
               f_cn[:, pack.C_NSTEPS].sum() / p_cn[:, pack.C_NSTEPS].sum(),
               f_cn[:, pack.C_NJEV].sum() / p_cn[:, pack.C_NJEV].sum(), f_cn[:, pack.C_NFAIL].sum(),
               p_cn[:, pack.C_NFAIL].sum()))
        assert np.all(f_ws[:, 1:, pack.W_STATUS] == 0), mode
        assert d.max() <= 1 and np.mean(d == 0) >= 0.97, mode
        assert np.median(psi_units) < 30.0, mode
        assert th_rel < 3e-3, mode
        assert rhs_ratio < 0.9, mode
