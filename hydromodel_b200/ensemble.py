"""
EnsembleSimulation -- many independent soil columns integrated on one GPU (or
on the GPUs of one box, one process per GPU).

It is the ensemble counterpart of the reference's single-column ``Simulation``
(code/src/simulation.py:21-713): same phases and names (``setupModel``,
``initial_conditions``, ``run``, ``saveResults``), same forcing table, but one
``input_parameters`` dict per column (or, for large ensembles, one base dict plus
arrays of the per-column values) and outputs with a leading column axis.  The
whole window loop of every column runs on the device through the C-ABI
(hydromodel_b200/capi.py).

Forcing.  ``data`` is the reference's four-column table; a list of tables gives
per-column forcing series (``series_of_col`` maps columns to tables), and
``precip_mult`` scales the precipitation of a column.

Noise.  In ``noise="tape"`` mode every column consumes host-generated standard
normals in exactly the order the reference's ``Simulation(name, seed=c)`` draws
them (spin-up draw, run draw, refresh draws: simulation.py:426,561,601), so a
column of the ensemble is comparable one-to-one with a reference run.  In
``noise="philox"`` mode the fields come from a counter-based generator on the
device keyed by (seed, global column id, row) -- statistically equivalent, no
per-column parity, independent of how the ensemble is sharded -- the mode for
ensembles too large to tape.

Multi-GPU.  ``EnsembleSimulation.from_global`` takes the description of the WHOLE
ensemble plus (rank, world) and keeps this rank's contiguous block of columns
(``sharding.column_range``); global column ids key the ensemble members and the
noise streams, so the job integrates the same columns however it is split.  The
only collective is the all-reduce of the ensemble statistics (``statistics``).
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import capi, pack
from .sharding import column_range


class EnsembleSimulation:

    def __init__(self, name=None, seed=0, device=0, noise="tape"):
        self.name = name if name is not None else "Ensemble_None"
        self.seed, self.device, self.noise = int(seed), int(device), noise
        if noise not in ("tape", "philox"):
            raise ValueError(" EnsembleSimulation: noise must be 'tape' or 'philox'.")
        self.mData, self.output = {}, {}
        self.ctx = None
        self.first_column = 0          # global id of local column 0 (multi-GPU sharding)
        self.n_global = None           # columns of the whole job (from_global)

    # ------------------------------------------------------------------ setup
    def setupModel(self, params_list, data, precip_mult=None, first_column=0, quiet=True,
                   series_of_col=None):
        """``params_list``: one parameter dict per column (all on the same well,
        i.e. the same vertical grid); ``data``: the forcing table shared by the
        columns, or a list of tables with ``series_of_col[n_cols]``;
        ``precip_mult``: optional per-column precipitation multipliers.

        Dicts that differ only in the soil / water-content / model keys of
        ``pack.ENSEMBLE_KEYS`` are packed together with array operations."""
        if not params_list:
            raise ValueError(" EnsembleSimulation: no columns given.")
        n = len(params_list)
        mult = np.ones(n) if precip_mult is None else np.asarray(precip_mult, dtype=float)
        tables = data if isinstance(data, (list, tuple)) else [data]
        par = stat = x = first = None
        for base, cols, idx in pack.split_params_list(params_list):
            cols["precip_mult"] = mult[idx]
            part = pack.pack_ensemble(base, tables[0], cols, quiet=quiet)
            if par is None:
                x, first = part["x"], part
                par = np.empty((n, pack.NPARAM))
                stat = np.empty((n, pack.NSTATIC, x.size))
            elif part["x"].shape != x.shape or not np.array_equal(part["x"], x):
                raise RuntimeError(" EnsembleSimulation: all columns must share one vertical grid.")
            par[idx], stat[idx] = part["par"], part["stat"]
        self._finish_setup(params_list[0], first, par, stat, tables, series_of_col, first_column)

    def setupEnsemble(self, base_params, data, columns, first_column=0, quiet=True, series_of_col=None):
        """Struct-of-arrays form for large ensembles: ``base_params`` is the
        parameter dict shared by all columns, ``columns`` maps the per-column keys
        of ``pack.ENSEMBLE_KEYS`` ("n", "a0", "psi_sat", "epsilon", "Theta_Max",
        "Theta_Min", "Theta_Residual", "model", "profile", "precip_mult") to arrays
        with one entry per column."""
        tables = data if isinstance(data, (list, tuple)) else [data]
        part = pack.pack_ensemble(base_params, tables[0], columns, quiet=quiet)
        self._finish_setup(base_params, part, part["par"], part["stat"], tables, series_of_col, first_column)

    @classmethod
    def from_global(cls, base_params, data, columns, n_total, rank=None, world=None, name=None,
                    seed=0, device=None, noise="philox", series_of_col=None):
        """This rank's share of a ``n_total``-column ensemble.  ``columns`` is
        either a dict of GLOBAL per-column arrays (length ``n_total``) or a callable
        ``columns(first, count) -> (base_params, columns)`` that builds the members
        ``first .. first + count - 1`` (e.g. ``synthetic.ensemble_columns``).
        ``rank`` / ``world`` default to the torch.distributed (or torchrun
        environment) values; ``device`` defaults to LOCAL_RANK."""
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        first, count, base_params, local, ser = cls.local_share(base_params, columns, n_total, rank, world,
                                                                 series_of_col)
        sim = cls(name=name, seed=seed, device=device, noise=noise)
        sim.setupEnsemble(base_params, data, local, first_column=first, series_of_col=ser)
        sim.n_global = int(n_total)
        return sim

    @staticmethod
    def local_share(base_params, columns, n_total, rank=None, world=None, series_of_col=None):
        """Host-side sharding of `from_global` (no device needed): this rank's block
        ``(first, count, base_params, local columns, local series map)``."""
        if rank is None:
            rank = int(os.environ.get("RANK", "0"))
        if world is None:
            world = int(os.environ.get("WORLD_SIZE", "1"))
        first, count = column_range(int(n_total), int(rank), int(world))
        if count < 1:
            raise ValueError(" EnsembleSimulation: this rank has no columns (n_total < world).")
        if callable(columns):
            base_params, local = columns(first, count)
        else:
            local = {k: np.asarray(v)[first:first + count] for k, v in columns.items()}
        ser = None if series_of_col is None else np.asarray(series_of_col)[first:first + count]
        return first, count, base_params, local, ser

    def _finish_setup(self, params0, base, par, stat, tables, series_of_col, first_column):
        m = self.mData
        x, n = base["x"], par.shape[0]
        m["z_grid"], m["n_cols"], m["dz"] = x, n, base["objs"]["dz"]
        m["par"], m["stat"] = np.ascontiguousarray(par), np.ascontiguousarray(stat)
        # forcing: one series, or several with a column -> series map
        series = [base["forcing"]] + [pack.pack_forcing(base["objs"], params0, t) for t in tables[1:]]
        dim_t = series[0]["dim_t"]
        if any(s["dim_t"] != dim_t for s in series):
            raise RuntimeError(" EnsembleSimulation: all forcing tables must have the same number of rows.")
        if any(s["interception"] != series[0]["interception"] or s["surface_evap"] != series[0]["surface_evap"]
               for s in series[1:]):
            # interception / surface evaporation are per-column scalars derived from the table
            # (simulation.py:347-352): re-derive them per column below
            pass
        f = {k: np.stack([s[k] for s in series]) for k in
             ("precipitation_cm", "wtd_idx", "daylight", "atm", "zWtd_cm")}
        f["time"], f["dim_t"], f["n_series"] = series[0]["time"], dim_t, len(series)
        if len(series) > 1:
            if series_of_col is None:
                raise ValueError(" EnsembleSimulation: several forcing tables need series_of_col.")
            ser = np.ascontiguousarray(series_of_col, dtype=np.int32)
            if ser.shape != (n,) or ser.min() < 0 or ser.max() >= len(series):
                raise ValueError(" EnsembleSimulation: series_of_col must map every column to a table.")
            m["par"][:, pack.P_INTERCEPT] = np.array([s["interception"] for s in series])[ser]
            m["par"][:, pack.P_SURF_EVAP] = np.array([s["surface_evap"] for s in series])[ser]
        else:
            ser = None
        m["series_of_col"] = ser
        m["forcing"], m["dim_t"] = f, dim_t
        m["objs"] = base["objs"]                      # objects of the base column (grid, layers, roots)
        m["groups"], m["n_groups"] = pack.fd_groups(x.size)
        self.first_column = int(first_column)
        m["initial_cond"] = None
        self._open()

    def _open(self):
        m = self.mData
        self.ctx = capi.Context(m["z_grid"].size, m["n_cols"], self.device)
        self.ctx.set_grid(m["z_grid"], m["groups"], m["n_groups"])
        self.ctx.set_columns(m["par"], m["stat"])
        f = m["forcing"]
        self.ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"],
                             series_of_col=m["series_of_col"])

    def _series(self):
        m = self.mData
        return np.zeros(m["n_cols"], dtype=np.int64) if m["series_of_col"] is None else m["series_of_col"]

    def _column_rng(self, local_col):
        # same stream as the reference's Simulation(name, seed): default_rng(SeedSequence(seed))
        return np.random.default_rng(np.random.SeedSequence(self.seed + self.first_column + local_col))

    def n_refresh_slots(self):
        """Tape slots a run needs: 1 + the largest number of rows on which a
        column refreshes its field (reference simulation.py:599; the row's
        precipitation includes the column's multiplier; skipped rows draw nothing)."""
        m = self.mData
        f, ser = m["forcing"], self._series()
        rows = np.arange(m["dim_t"])
        mult = m["par"][:, pack.P_PRECIP_MULT]
        most = 0
        for s in range(f["n_series"]):
            sel = ser == s
            if not np.any(sel):
                continue
            live = (rows > 0) & (f["wtd_idx"][s] >= 0)
            # the count grows with the multiplier: the largest one of the series decides
            refresh = ((mult[sel].max() * f["precipitation_cm"][s] > 0.5) | (rows % 48 == 0)) & live
            most = max(most, int(refresh.sum()))
        return 1 + most

    def set_initial_conditions(self, psi):
        """Shared (nx,) or per-column (n_cols, nx) pressure head: the ensemble
        form of the reference's IC_Filename (simulation.py:358-385)."""
        m = self.mData
        psi = np.asarray(psi, dtype=float)
        if psi.ndim == 1:
            psi = np.tile(psi, (m["n_cols"], 1))
        if psi.shape != (m["n_cols"], m["z_grid"].size):
            raise RuntimeError(" EnsembleSimulation: IC vector's dimensions do not match the spatial grid.")
        m["initial_cond"] = np.ascontiguousarray(psi)

    def _normals(self, rngs, shape):
        """One standard-normal block per column from that column's generator
        (NumPy releases the GIL while it fills an array)."""
        out = np.empty((len(rngs),) + tuple(shape))

        def fill(c):
            rngs[c].standard_normal(out=out[c])
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
            list(ex.map(fill, range(len(rngs))))
        return out

    def initial_conditions(self, max_iter=1500):
        """Spin-up of every column on the device (reference simulation.py:389-493).
        Returns the number of spin-up iterations per column."""
        m = self.mData
        n, nx = m["n_cols"], m["z_grid"].size
        # the saturated start psi_k = k * dz (simulation.py:414-417 through the inverse
        # retention curve, hydrological_model.py:91-115: the porosity profile maps to S_e = 1)
        psi0 = np.tile(np.arange(nx) * float(m["dz"]), (n, 1))
        self.ctx.set_state(psi0)
        f = m["forcing"]
        nrnd0 = None
        if self.noise == "tape":
            self._rngs = [self._column_rng(c) for c in range(n)]
            nrnd0 = self._normals(self._rngs, (nx,))
        else:
            self.ctx.set_noise_philox(self.seed, self.first_column)
        iters = self.ctx.spinup(f["zWtd_cm"][self._series(), 0], nrnd0, max_iter=max_iter)
        m["initial_cond"] = self.ctx.get_state()
        m["spinup_iters"] = iters
        return iters

    # -------------------------------------------------------------------- run
    def run(self, out_stride=30, windows=None):
        """Window loop of every column (reference simulation.py:495-672).
        ``self.output`` gets, with a leading column axis: per-row
        ``wtd_est_cm, abs_error, lateral_flow, status`` and, for the rows kept by
        ``out_stride``, ``psi_press, theta_vol, K_hrc, K_bkg, S_eff``."""
        m = self.mData
        if not m:
            raise RuntimeError(" EnsembleSimulation: Simulation data structure 'mData' is empty.")
        if m["initial_cond"] is None:
            self.initial_conditions()
        n, nx, T = m["n_cols"], m["z_grid"].size, m["dim_t"] if windows is None else int(windows)
        ctx = self.ctx
        if self.noise == "tape":
            rngs = getattr(self, "_rngs", None) or [self._column_rng(c) for c in range(n)]
            ctx.set_noise_tape(self._normals(rngs, (self.n_refresh_slots(), nx)))
        else:
            ctx.set_noise_philox(self.seed, self.first_column)
        ctx.set_output(out_stride)
        if out_stride > 0:
            # the result file: fail before the run, not after it, if the built-in writer could not hold it
            from .src import hdf5_io, hdf5_min
            if hdf5_io._h5py() is None:
                hdf5_min.check_writable([(n, ctx.n_output_rows, nx)])
        ctx.set_state(m["initial_cond"])
        ctx.reset_counters()
        ctx.run(0, T)
        wscal = ctx.get_wscal(rows=(0, T))
        z, f, ser = m["z_grid"], m["forcing"], self._series()
        wtd_idx = wscal[:, :, pack.W_WTD_EST].astype(np.int64)
        out = self.output
        out["wtd_est_cm"] = z[wtd_idx]
        out["abs_error"] = np.abs(f["zWtd_cm"][ser, :T] - out["wtd_est_cm"])
        out["lateral_flow"] = wscal[:, 1:, pack.W_LATERAL]
        out["transpiration"] = wscal[:, 1:, pack.W_TRANSP]
        out["status"] = wscal[:, :, pack.W_STATUS].astype(np.int8)
        if out_stride > 0:
            fields = ctx.get_fields()
            rows = [r for r in range(0, m["dim_t"], out_stride)]
            if (m["dim_t"] - 1) % out_stride:
                rows.append(m["dim_t"] - 1)
            out["output_rows"] = np.array(rows)
            out["psi_press"] = fields[:, :, pack.F_PSI]
            out["theta_vol"] = fields[:, :, pack.F_THETA]
            out["K_hrc"] = fields[:, :, pack.F_K]
            out["K_bkg"] = fields[:, :, pack.F_KBKG]
            out["S_eff"] = out["theta_vol"] / m["stat"][:, None, pack.S_PHI_NODE, :]
        out["counters"] = ctx.get_counters()
        m["final_state"] = ctx.get_state()
        return out

    # ------------------------------------------------------------- statistics
    def statistics(self, group=None, theta=True):
        """Ensemble statistics over ALL columns of the job: mean and standard
        deviation of the estimated water-table node index per row and (``theta``)
        of the water content at every output row and node, plus the number of
        columns that contributed per row.  The per-GPU moments are computed on the
        device; when torch.distributed is initialised they are all-reduced over the
        ranks (NCCL) -- the only collective of this path, once, after the window loop."""
        import torch
        from .sharding import reduce_moments
        m, ctx = self.mData, self.ctx
        T, nx = m["dim_t"], m["z_grid"].size
        dev = torch.device("cuda", self.device)
        n_out = ctx.n_output_rows if theta else 0
        wtd = torch.zeros(3, T, dtype=torch.float64, device=dev)
        th = torch.zeros(2, max(n_out, 1), nx, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        ctx.moments_dev(wtd[0].data_ptr(), wtd[1].data_ptr(), wtd[2].data_ptr(),
                        th[0].data_ptr() if n_out else None, th[1].data_ptr() if n_out else None)
        ctx.sync()
        mean, std, count = reduce_moments(wtd[0], wtd[1], wtd[2], group)
        res = {"wtd_node_mean": mean.cpu().numpy(), "wtd_node_std": std.cpu().numpy(),
               "count": count.cpu().numpy()}
        if n_out:
            n_cols = torch.full((1,), float(m["n_cols"]), dtype=torch.float64, device=dev)
            tmean, tstd, n_all = reduce_moments(n_cols, th[0], th[1], group)
            res["theta_mean"], res["theta_std"] = tmean.cpu().numpy(), tstd.cpu().numpy()
            res["n_columns"] = int(n_all.item())
        return res

    def wtd_statistics(self, group=None):
        """(mean, std, count) of the water-table node index per row."""
        s = self.statistics(group, theta=False)
        return s["wtd_node_mean"], s["wtd_node_std"], s["count"]

    # ------------------------------------------------------------------ output
    def saveResults(self, path=None):
        """Write ``self.output`` in the reference's HDF5 layout (root-level gzip
        datasets, simulation.py:697-707) with the extra leading column axis."""
        from .src.hdf5_io import save_datasets
        if not self.output:
            print(f" {type(self).__name__}: Simulation data structure 'output' is empty.")
            return None
        file_out = path or (self.name.strip().replace(" ", "_") + ".h5")
        print(f" Saving the results to: {file_out}")
        save_datasets(file_out, {k: np.asarray(v) for k, v in self.output.items()})
        return file_out

    def close(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None
