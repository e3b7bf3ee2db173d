"""Single-column simulation driver -- drop-in for the reference's ``Simulation``.

Same workflow and public names as code/src/simulation.py:21-746:

    sim = Simulation("Sim_01", seed)     sim.setupModel(params, water_data)
    sim.run()                            sim.saveResults()

``setupModel`` builds the same objects into ``mData`` (soil, theta, K, porosity,
tree_roots, hydro_model, forcing arrays, ...).  ``initial_conditions`` and ``run``
do not loop over windows on the host: the whole spin-up / window loop executes on
the GPU (hb_spinup / hb_run) with the noise fields drawn here, on the host, in
exactly the reference's RNG order, so a seeded run is comparable with the
reference window by window.  ``self.output`` has the reference's keys and shapes.
"""
import time
from pathlib import Path

import numpy as np
import pandas as pd
from numpy.random import default_rng, SeedSequence

from .. import capi, pack
from .hdf5_io import save_datasets, load_datasets
from .models.vanGenuchten import vanGenuchten
from .models.vrettas_fung import VrettasFung
from .richards_pde import RichardsPDE
from .utilities import find_wtd

BURN_IN = 1500          # reference simulation.py:420


class Simulation:
    __slots__ = ("name", "mData", "rng", "pde_model", "output", "_col", "_ctx")

    def __init__(self, name=None, seed=None):
        self.name = name if name is not None else "ID_None"
        self.mData = {}
        self.rng = default_rng(SeedSequence(seed)) if seed is not None else default_rng()
        self.pde_model = None
        self.output = {}
        self._col, self._ctx = None, None

    # ------------------------------------------------------------------ setup
    def setupModel(self, params, data):
        col = pack.pack_column(params, data, quiet=False)
        objs, forcing, m = col["objs"], col["forcing"], self.mData
        m["Well_No"] = params["Well_No"]
        for key in ("sat_cells", "layers", "dz", "z_grid", "soil", "theta", "K", "env_param",
                    "sim_flags", "porosity", "tree_roots", "LAI"):
            m[key] = objs[key]
        model_cls = VrettasFung if objs["model_id"] == 0 else vanGenuchten
        m["hydro_model"] = model_cls(objs["soil"], objs["porosity"], objs["K"], objs["theta"].res, objs["dz"])
        print(" Selected model: Vrettas-Fung" if objs["model_id"] == 0 else " Selected model: vanGenuchten")
        m["time"] = forcing["time"]
        m["zWtd_cm"] = forcing["zWtd_cm"]
        m["precipitation_cm"] = forcing["precipitation_cm"]
        m["interception"] = forcing["interception"]
        m["dim_t"] = forcing["dim_t"]
        m["atm"] = forcing["atm"]
        m["iPsi_50"] = col["par"][pack.P_IPSI50]
        m["surface_evap"] = forcing["surface_evap"]
        self._col = col
        self.pde_model = RichardsPDE(m)
        self._open()

        ic_file = params["IC_Filename"]
        if ic_file is None:
            m["initial_cond"] = self.initial_conditions()
        else:
            with open(Path(ic_file), "r") as fh:
                init_cond = np.array(pd.read_csv(fh, names=["IC"]).loc[:, "IC"])
            if m["z_grid"].shape != init_cond.shape:
                raise RuntimeError(f" {type(self).__name__}: IC vector's dimensions do not match the spatial grid.")
            print(f" IC vector was loaded successfully from: {ic_file}.")
            m["initial_cond"] = init_cond

    def _open(self):
        col = self._col
        x, f = col["x"], col["forcing"]
        grp, ng = pack.fd_groups(x.size)
        ctx = capi.Context(x.size, 1)
        ctx.set_grid(x, grp, ng)
        ctx.set_columns(col["par"][None], col["stat"][None])
        ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])
        self._ctx = ctx

    def initial_conditions(self):
        """Burn-in on the device: the unit window (0, 1) repeated with the first
        forcing row under SPINUP flags until the estimated water table is within
        2 dz of the first observation and the state is stationary."""
        m, ctx = self.mData, self._ctx
        m["sim_flags"]["SPINUP"] = True
        well_no = m["Well_No"]
        print(f" [Initial Conditions for Well no. {well_no}] Burn in period started ...")
        z = m["z_grid"]
        y0 = pack.saturated_initial_state(self._col["objs"])
        n_rnd = self.rng.standard_normal(z.size)
        ctx.set_state(y0[None])
        iters = int(ctx.spinup(m["zWtd_cm"][0], n_rnd[None], max_iter=BURN_IN)[0])
        y_end = ctx.get_state()[0]
        wtd_est = find_wtd(y_end >= m["soil"].psi_sat)
        abs_error = np.abs(m["zWtd_cm"][0] - z[wtd_est])
        if iters < BURN_IN or abs_error <= 2.0 * m["dz"]:
            print(f" [Initial Conditions for Well no. {well_no}] finished at [itr: {iters - 1}]"
                  f" with [abs(error): {abs_error}]")
        if iters >= BURN_IN:
            print(f" [Initial Conditions for Well no. {well_no}] finished at maximum number of iterations.")
        m["sim_flags"]["SPINUP"] = False
        m["spinup_iters"] = iters
        return y_end

    # -------------------------------------------------------------------- run
    def run(self):
        m, ctx = self.mData, self._ctx
        if not m:
            raise RuntimeError(f" {type(self).__name__}: Simulation data structure 'mData' is empty.")
        z, dim_t = m["z_grid"], m["dim_t"]
        dim_d = z.size
        f = self._col["forcing"]
        # noise in the reference's draw order: the run's field, then one field per
        # row with precipitation > 0.5 or row % 48 == 0 (simulation.py:561,599-602)
        rows = np.arange(dim_t)
        refresh = ((f["precipitation_cm"] > 0.5) | (rows % 48 == 0)) & (rows > 0) & (f["wtd_idx"] >= 0)
        tape = self.rng.standard_normal((1 + int(refresh.sum()), dim_d))
        ctx.set_noise_tape(tape[None])
        ctx.set_output(1)
        ctx.set_state(np.asarray(m["initial_cond"], dtype=float)[None])
        ctx.reset_counters()
        t0 = time.time()
        ctx.run(0, dim_t)
        wscal = ctx.get_wscal()[0]
        fields = ctx.get_fields()[0]
        print(" Elapsed time: {0:.2f} seconds.\n".format(time.time() - t0))
        done = wscal[:, pack.W_STATUS] != 2
        wtd_est = wscal[:, pack.W_WTD_EST].astype(int)
        abs_error = np.where(done, np.abs(m["zWtd_cm"] - z[wtd_est]), 0.0)
        for i in range(100, dim_t, 100):
            side = "-" if m["zWtd_cm"][i] < z[wtd_est[i]] else "+"
            print(" [Well No. {0}] {1}: MAE = {2:.2f} cm, [{3}]".format(
                m["Well_No"], i, np.mean(abs_error[0:i]), side))
        theta_vol = fields[:, pack.F_THETA]
        out = self.output
        out["K_hrc"] = fields[:, pack.F_K]
        out["K_bkg"] = fields[:, pack.F_KBKG]
        out["S_eff"] = theta_vol / m["porosity"]()[0]
        out["psi_press"] = fields[:, pack.F_PSI]
        out["theta_vol"] = theta_vol
        out["abs_error"] = abs_error
        out["wtd_est_cm"] = z[wtd_est]
        solved = done[1:]
        out["lateral_flow"] = wscal[1:, pack.W_LATERAL][solved]
        out["transpiration"] = wscal[1:, pack.W_TRANSP][solved]
        m["counters"] = dict(zip(capi.COUNTER_NAMES, ctx.get_counters()[0].tolist()))

    # ------------------------------------------------------------------ output
    def saveResults(self):
        if not self.output:
            print(f" {type(self).__name__}: Simulation data structure 'output' is empty.")
            return
        print(f" Saving the results to: {self.name}.h5")
        save_datasets(Path(self.name.strip().replace(" ", "_") + ".h5"), self.output)


def loadResults(filename=None):
    """Dictionary of the arrays stored by ``Simulation.saveResults``."""
    if filename is None:
        raise RuntimeError(" load_data: No input file is given.")
    return load_datasets(Path(filename))
