"""Quick throughput probe: C perturbed columns x T windows on one GPU."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hydromodel_b200 import capi, pack
from hydromodel_b200.synthetic import synthetic_forcing, default_params, ensemble_member
from helpers import SITE

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 61
forcing = synthetic_forcing(365)
base = default_params(SITE)
t0 = time.time()
pars, stats = [], []
uniq = min(C, 64)
for c in range(uniq):
    p, mult = ensemble_member(base, c, 1024)
    col = pack.pack_column(p, forcing)
    pars.append(col["par"]); stats.append(col["stat"])
par = np.stack([pars[c % uniq] for c in range(C)]); stat = np.stack([stats[c % uniq] for c in range(C)])
print("pack %.1fs" % (time.time() - t0))
f = col["forcing"]
g = np.load(os.path.join(ROOT, "tests", "golden", "windows_default.npz"))
grp, ng = pack.fd_groups(401)
ctx = capi.Context(401, C)
ctx.set_grid(col["x"], grp, ng)
ctx.set_columns(par, stat)
ctx.set_forcing(f["precipitation_cm"][:T], f["wtd_idx"][:T], f["daylight"][:T], f["atm"][:T])
ctx.set_output(30)
ctx.set_noise_philox(1234, 0)
if os.environ.get("HB_PROBE_FAST"):        # fast mode (never the default): 1 = continuation + 3-colour Jacobian, 2 = + f32 closure
    ctx.set_fast_mode(True, True, os.environ["HB_PROBE_FAST"] == "2")
for rep in range(3):
    ctx.set_state(np.tile(g["ic"], (C, 1)))
    ctx.reset_counters()
    ctx.sync()
    t0 = time.time()
    ctx.run(0, T)
    ctx.sync()
    dt = time.time() - t0
    cn = ctx.get_counters().sum(axis=0)
    print("rep %d: %.3f s  -> %.1f column-windows/s  %.2f column-years/s ; per window nfev %.1f njev %.2f nlu %.1f steps %.1f rhs %.1f fails %d" % (
        rep, dt, C * (T - 1) / dt, C * (T - 1) / dt / 365.0, cn[0] / cn[8], cn[1] / cn[8], cn[2] / cn[8], cn[3] / cn[8], cn[6] / cn[8], cn[5]))
pc = ctx.phase_cycles()
if pc[:8].sum() > 0:
    names = ["glue+top barrier", "closure phase", "assembly", "residual", "tri solve", "update+norm", "step ctl+predictor", "tri factor"]
    tot = pc[:8].sum()
    print("phase cycles of thread 0 (all columns, last rep): " + ", ".join(f"{n} {100*v/tot:.1f}%" for n, v in zip(names, pc[:8])))
    print("cycles per column-window: %.0f" % (tot / (C * (T - 1))))
    cn2 = ctx.get_counters().sum(axis=0)
    sub = ["loads", "fwd sweep", "scale+bwd", "PCR", "backsub"]
    print("tri_solve per call (cycles): " + ", ".join(f"{n} {pc[8+k]/max(cn2[7],1):.0f}" for k, n in enumerate(sub)), " | whole phase 4 per call: %.0f" % (pc[4] / max(cn2[7], 1)))
w = ctx.get_wscal()
print("wtd_est spread at end:", np.unique(w[:, -1, 0])[:10])
ctx.close()
