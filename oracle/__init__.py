"""
CPU oracle for the Richards time-stepping path -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper around oracle/hb_oracle.c (a plain-C restatement of the
reference algorithm, pinned against golden outputs of the unmodified reference:
see the header of hb_oracle.c and tests/test_oracle.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; hydromodel_b200/ never does.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

NCOUNTER, NWSCAL, NFIELD = 9, 4, 4


def build(force=False):
    """Compile oracle/hb_oracle.c with gcc (idempotent)."""
    src = os.path.join(_HERE, "hb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("HB_ORACLE_LIB")     # experiments only
        if not path:
            build()
            path = _LIB_PATH
        _lib = ct.CDLL(path)
        _lib.orc_solve_window_c.restype = ct.c_int
        _lib.orc_run_c.restype = ct.c_int
        _lib.orc_spinup_c.restype = ct.c_int
        _lib.orc_find_wtd.restype = ct.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_double))


def _i(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_int))


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def find_wtd(sat):
    s = np.ascontiguousarray(sat, dtype=np.uint8)
    return int(lib().orc_find_wtd(s.ctypes.data_as(ct.POINTER(ct.c_ubyte)), ct.c_int(s.size)))


def rhs(x, par, stat, psi, nrnd, precip, wtd, daylight=0, spinup=0, with_cond=False, atm=0.0,
        with_transpiration=False):
    """dpsi/dt and the lateral flow.  ``with_cond`` also returns, per node, the
    magnitude sum of the terms whose (cancelling) sum forms dpsi/dt -- the scale
    against which rounding-level differences have to be judged."""
    x, par, stat, psi, nrnd = _c(x), _c(par), _c(stat), _c(psi), _c(nrnd).copy()
    out = np.empty_like(psi)
    cond = np.empty_like(psi)
    lat = ct.c_double(0.0)
    transp = ct.c_double(0.0)
    lib().orc_rhs_c(ct.c_int(x.size), _d(x), _d(par), _d(stat), _d(psi), _d(nrnd),
                    ct.c_double(precip), ct.c_int(int(wtd)), ct.c_int(int(daylight)),
                    ct.c_int(int(spinup)), _d(out), ct.byref(lat), _d(cond), ct.c_double(atm),
                    ct.byref(transp))
    if with_transpiration:
        return out, lat.value, cond, transp.value
    if with_cond:
        return out, lat.value, cond
    return out, lat.value


def closure_nodes(x, par, stat, psi, nrnd):
    x, par, stat, psi, nrnd = _c(x), _c(par), _c(stat), _c(psi), _c(nrnd)
    out = np.empty((4, x.size))
    qinf = ct.c_double(0.0)
    lib().orc_nodes_c(ct.c_int(x.size), _d(x), _d(par), _d(stat), _d(psi), _d(nrnd),
                      _d(out), ct.byref(qinf))
    return out, qinf.value


def num_jac(x, groups, par, stat, psi, nrnd, precip, wtd, daylight=0, spinup=0, fac=None):
    x, par, stat, psi, nrnd = _c(x), _c(par), _c(stat), _c(psi), _c(nrnd).copy()
    groups = _c(groups, np.int32)
    n = x.size
    dl, d, du, fo = (np.empty(n) for _ in range(4))
    fin = _d(_c(fac)) if fac is not None else None
    lib().orc_num_jac_c(ct.c_int(n), _d(x), _i(groups), ct.c_int(int(groups.max()) + 1),
                        _d(par), _d(stat), _d(psi), _d(nrnd), ct.c_double(precip),
                        ct.c_int(int(wtd)), ct.c_int(int(daylight)), ct.c_int(int(spinup)),
                        fin, _d(dl), _d(d), _d(du), _d(fo))
    return dl, d, du, fo


def group_rows(x, groups, par, stat, psi, nrnd, precip, wtd, h, daylight=0, spinup=0, fused=False):
    """rows[j, q] = f(psi + h on the colour group of column j)[j - 1 + q], q = 0..2: the numbers
    num_jac reads (SCIPY common.py:385-443), by one evaluation per group (``fused=False``) or by
    the one-pass scheme of the device's jac_fused (``fused=True``)."""
    x, par, stat, psi, nrnd, h = _c(x), _c(par), _c(stat), _c(psi), _c(nrnd).copy(), _c(h)
    groups = _c(groups, np.int32)
    n = x.size
    rows = np.empty((n, 3))
    rc = lib().orc_group_rows_c(ct.c_int(n), _d(x), _i(groups), ct.c_int(int(groups.max()) + 1),
                                _d(par), _d(stat), _d(psi), _d(nrnd), ct.c_double(precip),
                                ct.c_int(int(wtd)), ct.c_int(int(daylight)), ct.c_int(int(spinup)),
                                _d(h), ct.c_int(int(bool(fused))), _d(rows))
    if rc != 0:
        raise ValueError("orc_group_rows_c: the root-uptake sink is active (one evaluation per group only)")
    return rows


def solve_window(x, groups, par, stat, t0, t1, y0, nrnd, precip, wtd, daylight=0,
                 spinup=0, rtol=1e-3, atol=1e-3, max_acc=4096, atm=0.0):
    """One RichardsPDE.solve.  ``nrnd`` is damped in place on failure (Q4).
    Returns (y_end, stats dict, status, lateral_flow, accepted times)."""
    x, par, stat, y0 = _c(x), _c(par), _c(stat), _c(y0)
    groups = _c(groups, np.int32)
    assert nrnd.dtype == np.float64 and nrnd.flags.c_contiguous
    n = x.size
    y_end = np.empty(n)
    stats = np.zeros(NCOUNTER, dtype=np.int64)
    lat = ct.c_double(0.0)
    t_acc = np.empty(max_acc)
    n_acc = ct.c_int(0)
    st = lib().orc_solve_window_c(
        ct.c_int(n), _d(x), _i(groups), ct.c_int(int(groups.max()) + 1), _d(par), _d(stat),
        ct.c_double(rtol), ct.c_double(atol), ct.c_double(t0), ct.c_double(t1), _d(y0),
        _d(nrnd), ct.c_double(precip), ct.c_int(int(wtd)), ct.c_int(int(daylight)),
        ct.c_int(int(spinup)), _d(y_end),
        stats.ctypes.data_as(ct.POINTER(ct.c_int64)), ct.byref(lat), _d(t_acc),
        ct.byref(n_acc), ct.c_int(max_acc), ct.c_double(atm), None)
    names = ("nfev", "njev", "nlu", "nsteps", "nreject", "nfail", "nrhs", "nnewton", "nwindows")
    return y_end, dict(zip(names, stats.tolist())), st, lat.value, t_acc[:n_acc.value].copy()


def run(x, groups, par, stat, precip, wtd_idx, daylight, atm, noise_tape, psi0,
        out_stride=0, rtol=1e-3, atol=1e-3):
    """Window loop of Simulation.run for one column (host noise tape)."""
    x, par, stat = _c(x), _c(par), _c(stat)
    groups = _c(groups, np.int32)
    precip, atm = _c(precip), _c(atm)
    wtd_idx, daylight = _c(wtd_idx, np.int32), _c(daylight, np.int32)
    tape = _c(noise_tape).copy()
    psi = _c(psi0).copy()
    n, T = x.size, precip.size
    wscal = np.zeros((T, NWSCAL))
    n_out = 0 if out_stride <= 0 else len({i for i in range(T) if i % out_stride == 0} | {T - 1})
    fields = np.zeros((max(n_out, 1), NFIELD, n))
    stats = np.zeros(NCOUNTER, dtype=np.int64)
    used = lib().orc_run_c(
        ct.c_int(n), _d(x), _i(groups), ct.c_int(int(groups.max()) + 1), _d(par), _d(stat),
        ct.c_double(rtol), ct.c_double(atol), ct.c_int(T), _d(precip), _i(wtd_idx),
        _i(daylight), _d(atm), _d(tape), ct.c_int(tape.shape[0]), _d(psi), _d(wscal),
        _d(fields) if n_out else None, ct.c_int(out_stride),
        stats.ctypes.data_as(ct.POINTER(ct.c_int64)))
    if used < 0:
        raise RuntimeError("oracle run: noise tape exhausted")
    return psi, wscal, fields[:n_out], stats, used


def spinup(x, groups, par, stat, precip0, wtd0, zwtd0_cm, nrnd, psi0, max_iter=1500,
           rtol=1e-3, atol=1e-3):
    x, par, stat = _c(x), _c(par), _c(stat)
    groups = _c(groups, np.int32)
    nr = _c(nrnd).copy()
    psi = _c(psi0).copy()
    stats = np.zeros(NCOUNTER, dtype=np.int64)
    it = lib().orc_spinup_c(
        ct.c_int(x.size), _d(x), _i(groups), ct.c_int(int(groups.max()) + 1), _d(par),
        _d(stat), ct.c_double(rtol), ct.c_double(atol), ct.c_double(precip0),
        ct.c_int(int(wtd0)), ct.c_double(zwtd0_cm), _d(nr), _d(psi), ct.c_int(max_iter),
        stats.ctypes.data_as(ct.POINTER(ct.c_int64)))
    return psi, it, stats, nr
