"""Tiny GPU exercise of every kernel path (used under compute-sanitizer)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hydromodel_b200 import capi, pack
from helpers import default_column

col = default_column()
f = col["forcing"]
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "windows_default.npz"))
grp, ng = pack.fd_groups(401)
ctx = capi.Context(401, 3)
ctx.set_grid(col["x"], grp, ng)
ctx.set_columns(np.tile(col["par"], (3, 1)), np.tile(col["stat"], (3, 1, 1)))
dydt, lat = ctx.rhs(0, g["ic"], g["n_rnd"][0], 0.3, 170)
print("rhs ok", float(np.abs(dydt).max()), lat)
nr = g["n_rnd"][0].copy()
y, st, status, lat = ctx.solve_window(0, 0.0, 1.0, g["y0"][0], nr, float(g["precip"][0]), int(g["wtd"][0]))
print("window ok", st, status, float(np.abs(y - g["y_end"][0]).max()))
T = 4
ctx.set_forcing(f["precipitation_cm"][:T], f["wtd_idx"][:T], f["daylight"][:T], f["atm"][:T])
ctx.set_output(2)
ctx.set_noise_tape(np.tile(np.stack([g["n_rnd"][0]] * 4)[None], (3, 1, 1)))
ctx.set_state(np.tile(g["ic"], (3, 1)))
ctx.run(0, T)
print("run ok", ctx.get_wscal()[0, :, 0], ctx.get_counters()[0])
ctx.close()
