"""Richards' equation on one column -- device-backed drop-in.

Keeps the reference's ``RichardsPDE`` surface (code/src/richards_pde.py:8-595):
construction from the ``m_data`` dictionary with the same validation,
``solve(t_span, y0, args_i)`` (one BDF solve of one forcing window, 5 attempts,
in-place noise damping, partial state on failure), ``__call__(t, y, args_i)``
(the method-of-lines right-hand side), the ``arg_out`` side channel and the
``midpoints`` helper.  Both ``solve`` and ``__call__`` execute on the GPU through
the C-ABI (hb_solve_window / hb_rhs); there is no host implementation of them.
"""
import numpy as np

from .. import capi, pack


class RichardsPDE:
    __slots__ = ("m_data", "x_mesh", "x_mid", "mid_i", "xzmp", "zxmp", "nx", "h_model",
                 "var_arg_out", "sim_flags", "_ctx", "_packed")

    def __init__(self, m_data=None):
        if m_data is None:
            raise ValueError(f" {type(self).__name__}: No input is given. The model cannot initialize.")
        self.var_arg_out = {"transpiration": 0.0, "lateral_flow": 0.0}
        self.m_data = m_data
        self.h_model = m_data["hydro_model"]
        self.x_mesh = m_data["z_grid"]
        self.sim_flags = m_data["sim_flags"]
        self.nx = self.x_mesh.size
        dx = np.diff(self.x_mesh, axis=0)
        if np.any(dx <= 0.0):
            raise RuntimeError(f" {type(self).__name__}: Space domain is not increasing.")
        self.x_mid = self.x_mesh[0:-1] + 0.5 * dx
        self.mid_i = np.arange(1, self.nx - 1)
        self.xzmp = np.zeros(self.nx)
        self.xzmp[1:] = self.x_mesh[1:] - self.x_mid
        self.zxmp = self.x_mid - self.x_mesh[0:-1]
        self._ctx, self._packed = None, None

    # -- device plumbing ------------------------------------------------------
    def _context(self):
        """One-column device context, created on first use from m_data."""
        if self._ctx is None:
            m = self.m_data
            objs = {"z_grid": m["z_grid"], "soil": m["soil"], "theta": m["theta"], "K": m["K"],
                    "porosity": m["porosity"], "tree_roots": m["tree_roots"], "layers": m["layers"],
                    "dz": m["dz"], "sim_flags": m["sim_flags"], "LAI": m.get("LAI", 0.0),
                    "model_id": 0 if type(m["hydro_model"]).__name__ == "VrettasFung" else 1}
            par, stat = pack.pack_static(objs, m.get("interception", 0.0), m.get("surface_evap", 0.0))
            grp, ng = pack.fd_groups(self.nx)
            ctx = capi.Context(self.nx, 1)
            ctx.set_grid(self.x_mesh, grp, ng)
            ctx.set_columns(par[None], stat[None])
            self._ctx, self._packed = ctx, (par, stat)
        return self._ctx

    @staticmethod
    def _daylight(args_i):
        stamp = args_i.get("time")
        return int(stamp is not None and 6 <= stamp.hour < 18)

    def _refresh_scalars(self):
        """interception / surface_evap may be set on m_data after construction."""
        ctx = self._context()
        par, stat = self._packed
        m = self.m_data
        want = (m.get("interception", 0.0), m.get("surface_evap", 0.0))
        if (par[pack.P_INTERCEPT], par[pack.P_SURF_EVAP]) != want:
            par[pack.P_INTERCEPT], par[pack.P_SURF_EVAP] = want
            ctx.set_columns(par[None], stat[None])
        return ctx

    # -- the reference's interface ---------------------------------------------
    def __call__(self, t, y, *args):
        a = args[0]
        ctx = self._refresh_scalars()
        dydt, lat, tr = ctx.rhs(0, np.atleast_1d(y), np.atleast_1d(a["n_rnd"]), a["precipitation"],
                                a["wtd"], self._daylight(a), int(bool(self.sim_flags["SPINUP"])),
                                atm=float(a.get("atm", 0.0)), with_transpiration=True)
        self.var_arg_out["lateral_flow"] = lat
        self.var_arg_out["transpiration"] = tr
        return dydt

    @property
    def arg_out(self):
        return self.var_arg_out

    def ic_fun(self, y0, *args):
        return y0

    def solve(self, t_span, y0, *args):
        a = args[0]
        ctx = self._refresh_scalars()
        n_rnd = a["n_rnd"]
        if not (isinstance(n_rnd, np.ndarray) and n_rnd.dtype == np.float64 and n_rnd.flags.c_contiguous):
            raise ValueError(f" {type(self).__name__}: n_rnd must be a contiguous float64 array.")
        y_end, stats, status, lat = ctx.solve_window(
            0, t_span[0], t_span[1], y0, n_rnd, a["precipitation"], a["wtd"],
            self._daylight(a), int(bool(self.sim_flags["SPINUP"])), atm=float(a.get("atm", 0.0)))
        if status != 0:
            print(f" {type(self).__name__}: The ODE solver failed with message: "
                  f"Required step size is less than spacing between numbers.")
        self.var_arg_out["lateral_flow"] = lat
        self.var_arg_out["transpiration"] = ctx.last_transpiration
        return y_end


def midpoints(x_left, fx_left, x_right, fx_right):
    """Mid-point values and central differences between two sets of points
    (reference richards_pde.py:542-595)."""
    x_left, fx_left, x_right, fx_right = np.atleast_1d(x_left, fx_left, x_right, fx_right)
    if fx_left.shape != fx_right.shape:
        raise RuntimeError(f" midpoints: Input 'fx' dimensions do not match."
                           f" {fx_left.shape} not equal to {fx_right.shape}.")
    if x_left.shape != x_right.shape:
        raise RuntimeError(f" midpoints: Input 'x' dimensions do not match."
                           f" {x_left.shape} not equal to {x_right.shape}. ")
    dx = x_right - x_left
    if fx_left.ndim == 2:
        dx = dx[:, None]
    return 0.5 * (fx_left + fx_right), (fx_right - fx_left) / dx
