"""Fused vs per-group Jacobian differences (HB_JAC_SEQ=1): same workload in two contexts, results must be bit-identical.
usage: python tools/jac_fuse_check.py [columns] [rows]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hydromodel_b200 import capi, pack
from hydromodel_b200.synthetic import synthetic_forcing, default_params, ensemble_member
from helpers import SITE

C = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
T = int(sys.argv[2]) if len(sys.argv) > 2 else 61
forcing = synthetic_forcing(365)
base = default_params(SITE)
pars, stats = [], []
uniq = min(C, 64)
for c in range(uniq):
    p, mult = ensemble_member(base, c, 1024)
    col = pack.pack_column(p, forcing)
    pars.append(col["par"]); stats.append(col["stat"])
par = np.stack([pars[c % uniq] for c in range(C)]); stat = np.stack([stats[c % uniq] for c in range(C)])
f = col["forcing"]
g = np.load(os.path.join(ROOT, "tests", "golden", "windows_default.npz"))
grp, ng = pack.fd_groups(401)
out = {}
for seq in (1, 0, 1, 0):
    if seq: os.environ["HB_JAC_SEQ"] = "1"
    else: os.environ.pop("HB_JAC_SEQ", None)
    ctx = capi.Context(401, C)
    ctx.set_grid(col["x"], grp, ng)
    ctx.set_columns(par, stat)
    ctx.set_forcing(f["precipitation_cm"][:T], f["wtd_idx"][:T], f["daylight"][:T], f["atm"][:T])
    ctx.set_output(30)
    ctx.set_noise_philox(1234, 0)
    ctx.set_state(np.tile(g["ic"], (C, 1)))
    ctx.reset_counters(); ctx.sync()
    t0 = time.time(); ctx.run(0, T); ctx.sync(); dt = time.time() - t0
    cn = ctx.get_counters()
    out[seq] = (ctx.get_state().copy(), ctx.get_wscal().copy(), cn.copy())
    s = cn.sum(axis=0)
    print("seq=%d: %.3f s -> %.1f column-windows/s ; per window nfev %.2f njev %.3f nlu %.2f steps %.2f rhs %.2f fails %d" % (
        seq, dt, C * (T - 1) / dt, s[0] / s[8], s[1] / s[8], s[2] / s[8], s[3] / s[8], s[6] / s[8], s[5]))
    ctx.close()
a, b = out[1], out[0]
print("state bit-identical:", np.array_equal(a[0].view(np.int64), b[0].view(np.int64)),
      " wscal bit-identical:", np.array_equal(a[1].view(np.int64), b[1].view(np.int64)),
      " counters identical:", np.array_equal(a[2], b[2]))
if not np.array_equal(a[0], b[0]):
    d = np.abs(a[0] - b[0]); print("max |dpsi|", d.max(), "columns differing", (d.max(axis=1) > 0).sum(), "of", C)
    bad = np.argwhere(a[2] != b[2]); print("counter diffs (first):", bad[:10].tolist())
