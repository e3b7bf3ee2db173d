"""Kernel time of the bench workload for the column block of a given rank (first column = rank * 1024)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from hydromodel_b200 import capi, pack
C, T = 1024, 365
g = np.load(os.path.join(ROOT, "tests", "golden", "windows_default.npz"))
IC = None
for rank in [int(a) for a in sys.argv[1:]] or [0, 1]:
    wl = bench.build_workload(C, rank * C, T, 8192)
    f = wl["forcing"]; nx = wl["x"].size
    if IC is None:
        IC = bench.default_ic(0, wl)[0] if os.environ.get("BENCH_IC") else g["ic"]
    tape = bench.noise_tape(C, rank * C, f, nx)
    grp, ng = pack.fd_groups(nx)
    ctx = capi.Context(nx, C)
    ctx.set_grid(wl["x"], grp, ng); ctx.set_columns(wl["par"], wl["stat"])
    ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])
    ctx.set_output(30)
    for rep in range(2):
        ctx.set_noise_tape(tape); ctx.set_state(np.tile(IC, (C, 1))); ctx.reset_counters(); ctx.sync()
        t0 = time.time(); ctx.run(0, T); ctx.sync(); dt = time.time() - t0
    cn = ctx.get_counters().sum(axis=0)
    print(f"rank {rank}: {dt*1e3:.1f} ms; per window nfev {cn[0]/cn[8]:.2f} rhs {cn[6]/cn[8]:.2f} steps {cn[3]/cn[8]:.2f} fails {cn[5]}", flush=True)
    per_col = ctx.get_counters()[:, 6]
    print("   rhs evals per column: min %d median %d max %d" % (per_col.min(), np.median(per_col), per_col.max()))
    top = np.argsort(per_col)[-3:][::-1]
    print("   heaviest columns:", [(int(k), int(per_col[k]), int(ctx.get_counters()[k, 5])) for k in top], "(local id, rhs evals, failed attempts)")
    ctx.close()
