import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hydromodel_b200 import capi, pack
nx = 401
ctx = capi.Context(nx, 1); grp, ng = pack.fd_groups(nx)
ctx.set_grid(np.arange(nx) * 5.0, grp, ng); ctx.set_columns(np.zeros((1, pack.NPARAM)), np.zeros((1, pack.NSTATIC, nx)))
rng = np.random.default_rng(0)
d = -rng.uniform(1, 500, nx); dl = np.r_[0, rng.uniform(.1, 200, nx - 1)]; du = np.r_[rng.uniform(.1, 200, nx - 1), 0]; b = rng.standard_normal(nx)
for k in range(3):
    x = ctx.tri_solve(dl, d, du, 0.01, b)
    print("call", k, "finite", bool(np.isfinite(x).all()))
