"""
Timing harness for the UNMODIFIED reference (oracle/_ref, see make_ref.py) --
BENCHMARK INFRASTRUCTURE ONLY: bench.py's `--impl reference` arm and its
cpu_baseline leg.  One ensemble member = one reference `Simulation(name, seed=c)`
(code/src/simulation.py:21-77) set up from that member's parameter dict with the
shared initial state handed over as an IC file (simulation.py:358-385, so the
19-61 s spin-up per column stays out of the timed loop) and integrated with
`Simulation.run` (simulation.py:495-672) over a bounded number of forcing rows.
A `multiprocessing` pool runs one member per host core; NumPy/BLAS threads are
pinned to 1.  The first call in every worker integrates one window untimed (Numba
JIT, simulation.py's imports).
"""
import contextlib
import io
import os
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_state = {}


def importable():
    """(ok, reason): can the reference run on this machine?"""
    if not os.path.isdir(os.path.join(REF_DIR, "src")):
        return False, "oracle/_ref is absent (run oracle/make_ref.py where /root/reference exists)"
    try:
        import numba  # noqa: F401
        import scipy  # noqa: F401
        import pandas  # noqa: F401
    except Exception as err:      # noqa: BLE001
        return False, f"reference dependency missing: {err}"
    return True, ""


def _init_worker(ic, site_path):
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ[k] = "1"
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    if "h5py" not in sys.modules:                 # only saveResults/loadResults touch it
        try:
            import h5py  # noqa: F401
        except Exception:                         # noqa: BLE001
            sys.modules["h5py"] = types.ModuleType("h5py")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    os.chdir(REF_DIR)
    from src.simulation import Simulation         # the reference's own class
    fh = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
    for v in ic:
        fh.write(repr(float(v)) + "\n")
    fh.close()
    _state.update(Simulation=Simulation, ic_file=fh.name, site=site_path, warm=False)


def _integrate(params, table, seed):
    sim = _state["Simulation"]("ref_member", seed=seed)
    p = dict(params, IC_Filename=_state["ic_file"], Site_Information=_state["site"])
    with contextlib.redirect_stdout(io.StringIO()):
        sim.setupModel(p, table)
        t0 = time.perf_counter()
        sim.run()
        dt = time.perf_counter() - t0
    return dt, sim.output["psi_press"][-1].copy()


def _member(job):
    params, table, seed, mult = job
    table = table.copy()
    if mult != 1.0:                               # ensemble precipitation multiplier
        table["Precipitation_cm"] = table["Precipitation_cm"] * mult
    if not _state["warm"]:
        _integrate(params, table.iloc[:2].reset_index(drop=True), seed)
        _state["warm"] = True
    dt, psi_end = _integrate(params, table, seed)
    return dt, psi_end


class ReferencePool:
    """Worker pool with the reference imported and JIT-warmed once per process."""

    def __init__(self, ic, site_path, procs=None):
        import multiprocessing as mp
        self.procs = procs or (os.cpu_count() or 1)
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_init_worker,
                                                 initargs=(np.asarray(ic, dtype=float), site_path))

    def run(self, members, table, n_rows):
        """``members``: list of (params, seed, precip_mult).  Integrates rows
        1 .. n_rows-1 of ``table`` for every member; returns (wall seconds,
        windows integrated, summed in-worker run() seconds, final states)."""
        tab = table.iloc[:n_rows].reset_index(drop=True)
        jobs = [(p, tab, seed, mult) for p, seed, mult in members]
        t0 = time.perf_counter()
        res = self.pool.map(_member, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return wall, len(members) * (n_rows - 1), float(sum(r[0] for r in res)), [r[1] for r in res]

    def warm(self, members, table):
        """Untimed: one window per worker so that every process has compiled."""
        tab = table.iloc[:2].reset_index(drop=True)
        jobs = [(members[k % len(members)][0], tab, members[k % len(members)][1], 1.0)
                for k in range(2 * self.procs)]
        self.pool.map(_member, jobs, chunksize=1)

    def close(self):
        self.pool.close()
        self.pool.join()
