"""Two builds of the library on the same workload: are the results bit-identical, and how fast is each?
usage: python tools/ab_bits.py <libA.so|-> <libB.so|-> [columns] [rows]      ("-" = the shipped library)"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, time
import numpy as np
ROOT = sys.argv[1]; out = sys.argv[2]; C = int(sys.argv[3]); T = int(sys.argv[4])
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hydromodel_b200 import capi, pack
from hydromodel_b200.synthetic import synthetic_forcing, default_params, ensemble_columns
from helpers import SITE
table = synthetic_forcing(365).iloc[:T].reset_index(drop=True)
base, cols = ensemble_columns(default_params(SITE), 0, min(C, 256), mixed=True, perturb_theta=True)
ens = pack.pack_ensemble(base, table, cols)
rep = (C + 255) // 256
par = np.tile(ens["par"], (rep, 1))[:C]; stat = np.tile(ens["stat"], (rep, 1, 1))[:C]
g = np.load(os.path.join(ROOT, "tests", "golden", "windows_default.npz"))
grp, ng = pack.fd_groups(401)
ctx = capi.Context(401, C)
ctx.set_grid(ens["x"], grp, ng)
ctx.set_columns(par, stat)
f = ens["forcing"]
ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])
ctx.set_output(10)
ctx.set_noise_philox(99, 0)
best = 1e9
for rep_ in range(2):
    ctx.set_state(np.tile(g["ic"], (C, 1))); ctx.reset_counters(); ctx.sync()
    t0 = time.time(); ctx.run(0, T); ctx.sync(); best = min(best, time.time() - t0)
np.savez(out, state=ctx.get_state(), wscal=ctx.get_wscal(), fields=ctx.get_fields(), counters=ctx.get_counters(), dt=best)
ctx.close()
'''
a, b = sys.argv[1], sys.argv[2]
C = sys.argv[3] if len(sys.argv) > 3 else "1184"
T = sys.argv[4] if len(sys.argv) > 4 else "41"
res = []
for lib in (a, b):
    env = dict(os.environ)
    env.pop("HB_LIB_PATH", None)
    if lib != "-":
        env["HB_LIB_PATH"] = os.path.abspath(lib)
    out = tempfile.mktemp(suffix=".npz")
    subprocess.run([sys.executable, "-c", CHILD, ROOT, out, C, T], env=env, check=True)
    res.append(np.load(out))
ra, rb = res
for k in ("state", "wscal", "fields", "counters"):
    print(k, "bit-identical:", np.array_equal(ra[k], rb[k], equal_nan=True))
n = int(C) * (int(T) - 1)
print("A %s: %.1f column-windows/s   B %s: %.1f column-windows/s" % (a, n / float(ra["dt"]), b, n / float(rb["dt"])))
