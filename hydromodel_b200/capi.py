"""
ctypes binding of libhydrob200.so (C-ABI: include/hydrob200.h).

This is the only door between the Python host layer and the device: NumPy
arrays in, NumPy arrays out, errors raised as ``HydroB200Error`` with the
library's message.  There is no CPU fallback -- if the library cannot be loaded
or no CUDA device is present the calls fail loudly.
"""
import ctypes as ct
import os

import numpy as np

from . import build as _build
from .pack import NPARAM, NSTATIC, NWSCAL, NFIELD, NCOUNTER

FAST_CONT, FAST_JAC3, FAST_F32 = 1, 2, 4          # include/hb_layout.h HB_FAST_*

_lib = None

_D = ct.POINTER(ct.c_double)
_I = ct.POINTER(ct.c_int32)
_L = ct.POINTER(ct.c_int64)

EXPORTS = {
    # name: (restype, argtypes)
    "hb_last_error": (ct.c_char_p, []),
    "hb_version": (ct.c_int, []),
    "hb_create": (ct.c_void_p, [ct.c_int, ct.c_int, ct.c_int]),
    "hb_destroy": (ct.c_int, [ct.c_void_p]),
    "hb_max_nx": (ct.c_int, []),
    "hb_set_stream": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
    "hb_sync": (ct.c_int, [ct.c_void_p]),
    "hb_set_grid": (ct.c_int, [ct.c_void_p, _D, _I, ct.c_int]),
    "hb_set_tolerances": (ct.c_int, [ct.c_void_p, ct.c_double, ct.c_double]),
    "hb_set_columns": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, _D, _D]),
    "hb_set_state": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, _D]),
    "hb_get_state": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, _D]),
    "hb_set_forcing": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, _D, _I, _I, _D, _I]),
    "hb_set_noise_tape": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, _D]),
    "hb_set_noise_philox": (ct.c_int, [ct.c_void_p, ct.c_uint64, ct.c_int64]),
    "hb_set_output": (ct.c_int, [ct.c_void_p, ct.c_int]),
    "hb_n_output_rows": (ct.c_int, [ct.c_void_p]),
    "hb_spinup": (ct.c_int, [ct.c_void_p, ct.c_int, _D, _D, _I]),
    "hb_run": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int]),
    "hb_set_schedule": (ct.c_int, [ct.c_void_p, ct.c_int]),
    "hb_set_fast_mode": (ct.c_int, [ct.c_void_p, ct.c_int]),
    "hb_get_wscal": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, _D]),
    "hb_get_wscal_rows": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, _D]),
    "hb_get_fields": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, _D]),
    "hb_get_fields_rows": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, _D]),
    "hb_get_counters": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, _L]),
    "hb_reset_counters": (ct.c_int, [ct.c_void_p]),
    "hb_moments_dev": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p,
                                  ct.c_void_p]),
    "hb_launch_count": (ct.c_int64, [ct.c_void_p]),
    "hb_measure_fp64_peak": (ct.c_double, [ct.c_void_p]),
    "hb_get_phase_cycles": (ct.c_int, [ct.c_void_p, _L]),
    "hb_rhs": (ct.c_int, [ct.c_void_p, ct.c_int, _D, _D, ct.c_double, ct.c_int, ct.c_int,
                          ct.c_int, ct.c_double, _D, _D, _D]),
    "hb_closure_nodes": (ct.c_int, [ct.c_void_p, ct.c_int, _D, _D, _D]),
    "hb_num_jac": (ct.c_int, [ct.c_void_p, ct.c_int, _D, _D, ct.c_double, ct.c_int, ct.c_int,
                              ct.c_int, _D, _D, _D, _D, _D]),
    "hb_tri_solve": (ct.c_int, [ct.c_void_p, _D, _D, _D, ct.c_double, _D, _D]),
    "hb_find_wtd": (ct.c_int, [ct.c_void_p, ct.c_int, _D, _I]),
    "hb_philox_normals": (ct.c_int, [ct.c_void_p, ct.c_uint64, ct.c_int64, ct.c_int, ct.c_int, _D]),
    "hb_solve_window": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_double, ct.c_double, _D, _D,
                                   ct.c_double, ct.c_int, ct.c_int, ct.c_int, ct.c_double, _D, _L,
                                   _D, _D, ct.POINTER(ct.c_int), _D]),
}

COUNTER_NAMES = ("nfev", "njev", "nlu", "nsteps", "nreject", "nfail", "nrhs", "nnewton",
                 "nwindows")


class HydroB200Error(RuntimeError):
    pass


def library_path():
    return _build.LIB


def load():
    """Load (building first if the sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    try:
        # HB_LIB_PATH: an alternative build of the same library (A/B experiments)
        path = os.environ.get("HB_LIB_PATH") or _build.build_library()
    except RuntimeError:
        if not os.path.exists(path):
            raise HydroB200Error(
                "libhydrob200.so is missing and cannot be built (no nvcc); "
                "there is no CPU fallback") from None
    lib = ct.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_D)


def _i(a):
    return a.ctypes.data_as(_I)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """One ensemble of columns on one GPU."""

    def __init__(self, nx, n_cols, device=0):
        self.lib = load()
        self.nx, self.n_cols, self.device = int(nx), int(n_cols), int(device)
        self.T = 0
        self._h = self.lib.hb_create(self.device, self.nx, self.n_cols)
        if not self._h:
            raise HydroB200Error(self.lib.hb_last_error().decode())

    # -- plumbing -----------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise HydroB200Error(self.lib.hb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.hb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:          # noqa: BLE001 - interpreter shutdown
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- setup --------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        self._ck(self.lib.hb_set_stream(self._h, ct.c_void_p(cuda_stream_handle)))

    def sync(self):
        self._ck(self.lib.hb_sync(self._h))

    def set_grid(self, x, groups, n_groups=None):
        x, groups = _f64(x), _i32(groups)
        assert x.size == self.nx and groups.size == self.nx
        ng = int(groups.max()) + 1 if n_groups is None else int(n_groups)
        self._ck(self.lib.hb_set_grid(self._h, _d(x), _i(groups), ng))

    def set_tolerances(self, rtol, atol):
        self._ck(self.lib.hb_set_tolerances(self._h, float(rtol), float(atol)))

    def set_columns(self, par, stat, col0=0):
        par = _f64(par).reshape(-1, NPARAM)
        stat = _f64(stat).reshape(par.shape[0], NSTATIC, self.nx)
        self._ck(self.lib.hb_set_columns(self._h, col0, par.shape[0], _d(par), _d(stat)))

    def set_state(self, psi, col0=0):
        psi = _f64(psi).reshape(-1, self.nx)
        self._ck(self.lib.hb_set_state(self._h, col0, psi.shape[0], _d(psi)))

    @staticmethod
    def _out(out, shape):
        """Caller-provided result buffer (e.g. pinned host memory) or a fresh array."""
        if out is None:
            return np.empty(shape)
        if out.shape != tuple(shape) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float64 array of shape {tuple(shape)}")
        return out

    def get_state(self, col0=0, ncol=None, out=None):
        ncol = self.n_cols - col0 if ncol is None else ncol
        out = self._out(out, (ncol, self.nx))
        self._ck(self.lib.hb_get_state(self._h, col0, ncol, _d(out)))
        return out

    def set_forcing(self, precip, wtd_idx, daylight, atm=None, series_of_col=None):
        precip = np.atleast_2d(_f64(precip))
        wtd_idx = np.atleast_2d(_i32(wtd_idx))
        daylight = np.atleast_2d(_i32(daylight))
        n_series, T = precip.shape
        assert wtd_idx.shape == precip.shape and daylight.shape == precip.shape
        atm_p = None
        if atm is not None:
            atm = np.atleast_2d(_f64(atm))
            atm_p = _d(atm)
        ser_p = None
        if series_of_col is not None:
            series_of_col = _i32(series_of_col)
            assert series_of_col.size == self.n_cols
            ser_p = _i(series_of_col)
        self._ck(self.lib.hb_set_forcing(self._h, n_series, T, _d(precip), _i(wtd_idx),
                                         _i(daylight), atm_p, ser_p))
        self.T = T

    def set_noise_tape(self, tape, col0=0):
        tape = _f64(tape)
        if tape.ndim == 2:
            tape = tape[None]
        ncol, n_slots, nx = tape.shape
        assert nx == self.nx
        self._ck(self.lib.hb_set_noise_tape(self._h, col0, ncol, n_slots, _d(tape)))

    def set_noise_philox(self, seed, first_global_column=0):
        self._ck(self.lib.hb_set_noise_philox(self._h, int(seed), int(first_global_column)))

    def set_output(self, out_stride):
        self._ck(self.lib.hb_set_output(self._h, int(out_stride)))

    @property
    def n_output_rows(self):
        return self.lib.hb_n_output_rows(self._h)

    # -- the path ---------------------------------------------------------------
    def spinup(self, zwtd0_cm, nrnd0=None, max_iter=1500, want_iters=True):
        zw = _f64(np.broadcast_to(zwtd0_cm, (self.n_cols,)))
        nr_p = None
        if nrnd0 is not None:
            nrnd0 = _f64(nrnd0).reshape(self.n_cols, self.nx)
            nr_p = _d(nrnd0)
        iters = np.zeros(self.n_cols, dtype=np.int32)
        self._ck(self.lib.hb_spinup(self._h, int(max_iter), nr_p, _d(zw),
                                    _i(iters) if want_iters else None))
        return iters

    def run(self, w0=0, w1=None):
        self._ck(self.lib.hb_run(self._h, int(w0), int(self.T if w1 is None else w1)))

    def set_fast_mode(self, cont=True, jac3=True, f32=False):
        """Fast mode of ``run`` (never the default): see HB_FAST_* in include/hb_layout.h."""
        self._ck(self.lib.hb_set_fast_mode(self._h, (FAST_CONT if cont else 0) | (FAST_JAC3 if jac3 else 0) |
                                           (FAST_F32 if f32 else 0)))

    def set_schedule(self, adaptive=True):
        self._ck(self.lib.hb_set_schedule(self._h, int(bool(adaptive))))

    def get_wscal(self, col0=0, ncol=None, out=None, rows=None):
        """Per-row scalars [ncol][T][NWSCAL]; ``rows=(row0, row1)`` reads that slice only."""
        ncol = self.n_cols - col0 if ncol is None else ncol
        r0, r1 = (0, self.T) if rows is None else (int(rows[0]), int(rows[1]))
        out = self._out(out, (ncol, r1 - r0, NWSCAL))
        self._ck(self.lib.hb_get_wscal_rows(self._h, col0, ncol, r0, r1 - r0, _d(out)))
        return out

    def get_fields(self, col0=0, ncol=None, out=None, out_rows=None):
        """Fields [ncol][n_out][NFIELD][nx]; ``out_rows=(o0, o1)`` reads those output rows only."""
        ncol = self.n_cols - col0 if ncol is None else ncol
        o0, o1 = (0, self.n_output_rows) if out_rows is None else (int(out_rows[0]), int(out_rows[1]))
        out = self._out(out, (ncol, o1 - o0, NFIELD, self.nx))
        self._ck(self.lib.hb_get_fields_rows(self._h, col0, ncol, o0, o1 - o0, _d(out)))
        return out

    def get_counters(self, col0=0, ncol=None):
        ncol = self.n_cols - col0 if ncol is None else ncol
        out = np.zeros((ncol, NCOUNTER), dtype=np.int64)
        self._ck(self.lib.hb_get_counters(self._h, col0, ncol, out.ctypes.data_as(_L)))
        return out

    def reset_counters(self):
        self._ck(self.lib.hb_reset_counters(self._h))

    def moments_dev(self, count_ptr, wtd_sum_ptr, wtd_sumsq_ptr, theta_sum_ptr=None, theta_sumsq_ptr=None):
        """Ensemble moments into DEVICE buffers (raw pointers): count / sum / sum of squares of
        the water-table node index per row [T], optionally sum / sum of squares of theta
        [n_out][nx]."""
        self._ck(self.lib.hb_moments_dev(self._h, ct.c_void_p(count_ptr), ct.c_void_p(wtd_sum_ptr),
                                         ct.c_void_p(wtd_sumsq_ptr), ct.c_void_p(theta_sum_ptr),
                                         ct.c_void_p(theta_sumsq_ptr)))

    @property
    def launch_count(self):
        return int(self.lib.hb_launch_count(self._h))

    def phase_cycles(self):
        out = np.zeros(16, dtype=np.int64)
        self._ck(self.lib.hb_get_phase_cycles(self._h, out.ctypes.data_as(_L)))
        return out

    def measure_fp64_peak(self):
        """TFLOP/s of a register-resident DFMA loop (roofline denominator)."""
        v = float(self.lib.hb_measure_fp64_peak(self._h))
        if v <= 0:
            raise HydroB200Error(self.lib.hb_last_error().decode())
        return v

    # -- test hooks ---------------------------------------------------------------
    def rhs(self, col, psi, nrnd, precip, wtd, daylight=0, spinup=0, atm=0.0, with_transpiration=False):
        psi, nrnd = _f64(psi), _f64(nrnd)
        out = np.empty(self.nx)
        lat, tr = ct.c_double(0.0), ct.c_double(0.0)
        self._ck(self.lib.hb_rhs(self._h, col, _d(psi), _d(nrnd), float(precip), int(wtd),
                                 int(daylight), int(spinup), float(atm), _d(out), ct.byref(lat),
                                 ct.byref(tr)))
        if with_transpiration:
            return out, lat.value, tr.value
        return out, lat.value

    def closure_nodes(self, col, psi, nrnd):
        psi, nrnd = _f64(psi), _f64(nrnd)
        out = np.empty((4, self.nx))
        self._ck(self.lib.hb_closure_nodes(self._h, col, _d(psi), _d(nrnd), _d(out)))
        return out

    def num_jac(self, col, psi, nrnd, precip, wtd, daylight=0, spinup=0, fac=None):
        psi, nrnd = _f64(psi), _f64(nrnd)
        dl, d, du, fo = (np.empty(self.nx) for _ in range(4))
        fin = None
        if fac is not None:
            fac = _f64(fac)
            fin = _d(fac)
        self._ck(self.lib.hb_num_jac(self._h, col, _d(psi), _d(nrnd), float(precip), int(wtd),
                                     int(daylight), int(spinup), fin, _d(dl), _d(d), _d(du), _d(fo)))
        return dl, d, du, fo

    def tri_solve(self, dl, d, du, c, b):
        dl, d, du, b = _f64(dl), _f64(d), _f64(du), _f64(b)
        x = np.empty(self.nx)
        self._ck(self.lib.hb_tri_solve(self._h, _d(dl), _d(d), _d(du), float(c), _d(b), _d(x)))
        return x

    def find_wtd(self, col, psi):
        psi = _f64(psi)
        assert psi.size == self.nx
        w = np.zeros(1, dtype=np.int32)
        self._ck(self.lib.hb_find_wtd(self._h, int(col), _d(psi), _i(w)))
        return int(w[0])

    def philox_normals(self, seed, global_column, window, n=None):
        out = np.empty(self.nx if n is None else int(n))
        self._ck(self.lib.hb_philox_normals(self._h, int(seed), int(global_column), int(window),
                                            out.size, _d(out)))
        return out

    def solve_window(self, col, t0, t1, y0, nrnd, precip, wtd, daylight=0, spinup=0, atm=0.0):
        """Returns (y_end, stats dict, status, lateral_flow); ``nrnd`` (float64,
        contiguous) is updated in place when the solver damped it.  The
        transpiration of the last RHS evaluation is left in ``self.last_transpiration``."""
        y0 = _f64(y0)
        assert nrnd.dtype == np.float64 and nrnd.flags.c_contiguous
        y_end = np.empty(self.nx)
        stats = np.zeros(NCOUNTER, dtype=np.int64)
        lat, tr = ct.c_double(0.0), ct.c_double(0.0)
        status = ct.c_int(0)
        t_acc = np.zeros(self.nx)
        self._ck(self.lib.hb_solve_window(self._h, col, float(t0), float(t1), _d(y0), _d(nrnd),
                                          float(precip), int(wtd), int(daylight), int(spinup),
                                          float(atm), _d(y_end), stats.ctypes.data_as(_L),
                                          ct.byref(lat), ct.byref(tr), ct.byref(status), _d(t_acc)))
        self.last_transpiration = tr.value
        # accepted step times of the last attempt (sol.t[1:] of scipy), for decision-level parity
        self.last_t_accepted = t_acc[:min(int(stats[3]), self.nx)].copy()
        return y_end, dict(zip(COUNTER_NAMES, stats.tolist())), status.value, lat.value
