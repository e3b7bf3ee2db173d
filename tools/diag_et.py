import sys, os, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from hydromodel_b200 import capi, pack
from hydromodel_b200.synthetic import synthetic_forcing, default_params
from helpers import SITE, tol_scaled
g = np.load(os.path.join(ROOT, "tests/golden/windows_default.npz"))
col = pack.pack_column(default_params(SITE), synthetic_forcing(96, half_hourly=True))
ctx = capi.Context(401, 1); grp, ng = pack.fd_groups(401)
ctx.set_grid(col["x"], grp, ng); ctx.set_columns(col["par"][None], col["stat"][None])
f = col["forcing"]; y = g["ic"].copy()
for row in (1, 2, 12, 13, 14, 15, 20, 30, 36, 37, 40):
    nr_d, nr_o = g["n_rnd"][row].copy(), g["n_rnd"][row].copy()
    day, atm = int(f["daylight"][row]), float(f["atm"][row])
    precip, wtd = float(f["precipitation_cm"][row]), int(f["wtd_idx"][row])
    t0 = time.time(); y_d, st_d, status_d, lat_d = ctx.solve_window(0, row - 1, row, y, nr_d, precip, wtd, day, 0, atm=atm); td = time.time() - t0
    t0 = time.time(); y_o, st_o, status_o, lat_o, _ = oracle.solve_window(col["x"], grp, col["par"], col["stat"], row - 1, row, y, nr_o, precip, wtd, day, 0, atm=atm); to = time.time() - t0
    print(f"row {row:2d} day {day} atm {atm:.4f} | dev status {status_d} nfev {st_d['nfev']:6d} nfail {st_d['nfail']} steps {st_d['nsteps']:5d} {td:.2f}s tr {ctx.last_transpiration:.3e} | orc status {status_o} nfev {st_o['nfev']:6d} nfail {st_o['nfail']} steps {st_o['nsteps']:5d} {to:.2f}s | err {tol_scaled(y_d, y_o):.3g}")
    y = y_o
