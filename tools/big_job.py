"""Large-ensemble smoke run (BASELINE configs[2] shape, scaled): C columns x T rows, device noise."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hydromodel_b200 import capi, pack
from hydromodel_b200.synthetic import synthetic_forcing, default_params, ensemble_member
from helpers import SITE
C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 365
forcing = synthetic_forcing(T)
base = default_params(SITE)
uniq = min(C, 128)
cols = [pack.pack_column(*(lambda pm: (pm[0], forcing))(ensemble_member(base, c, 65536))) for c in range(uniq)]
par = np.stack([cols[c % uniq]["par"] for c in range(C)]); stat = np.stack([cols[c % uniq]["stat"] for c in range(C)])
f = cols[0]["forcing"]
g = np.load(os.path.join(ROOT, "tests", "golden", "windows_default.npz"))
grp, ng = pack.fd_groups(401)
ctx = capi.Context(401, C)
ctx.set_grid(cols[0]["x"], grp, ng); ctx.set_columns(par, stat)
ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])
ctx.set_output(stride); ctx.set_noise_philox(2024, 0)
ctx.set_state(np.tile(g["ic"], (C, 1))); ctx.reset_counters(); ctx.sync()
t0 = time.time(); ctx.run(0, T); ctx.sync(); dt = time.time() - t0
cn = ctx.get_counters(); w = ctx.get_wscal()
print(f"{C} columns x {T} rows: {dt:.2f} s, {C * (T - 1) / dt:.0f} column-windows/s, {C * (T - 1) / dt / 365:.1f} column-years/s")
print("windows per column:", np.unique(cn[:, pack.C_NWINDOWS]), "failed attempts:", int(cn[:, pack.C_NFAIL].sum()),
      "rows with status != 0:", int((w[:, 1:, pack.W_STATUS] != 0).sum()), "finite:", bool(np.isfinite(ctx.get_state()).all()))
same = [np.array_equal(w[c], w[c % uniq]) for c in range(uniq, min(C, 4 * uniq))]
print("replicated members differ only by their noise stream:", not all(same))
ctx.close()
