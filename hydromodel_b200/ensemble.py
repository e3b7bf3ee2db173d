"""
EnsembleSimulation -- many independent soil columns integrated on one GPU (or
on the GPUs of one box, one process per GPU).

It is the ensemble counterpart of the reference's single-column ``Simulation``
(code/src/simulation.py:21-713): same phases and names (``setupModel``,
``initial_conditions``, ``run``, ``saveResults``), same forcing table, but a
list of parameter dictionaries (one ``input_parameters`` dict per column) and
outputs with a leading column axis.  The whole window loop of every column runs
on the device through the C-ABI (hydromodel_b200/capi.py).

Noise.  In ``noise="tape"`` mode every column consumes host-generated standard
normals in exactly the order the reference's ``Simulation(name, seed=c)`` draws
them (spin-up draw, run draw, refresh draws: simulation.py:426,561,601), so a
column of the ensemble is comparable one-to-one with a reference run.  In
``noise="philox"`` mode the fields come from a counter-based generator on the
device (statistically equivalent, no per-column parity) -- the mode for
ensembles too large to tape.
"""
import numpy as np

from . import capi, pack


class EnsembleSimulation:

    def __init__(self, name=None, seed=0, device=0, noise="tape"):
        self.name = name if name is not None else "Ensemble_None"
        self.seed, self.device, self.noise = int(seed), int(device), noise
        if noise not in ("tape", "philox"):
            raise ValueError(" EnsembleSimulation: noise must be 'tape' or 'philox'.")
        self.mData, self.output = {}, {}
        self.ctx = None
        self.first_column = 0          # global id of local column 0 (multi-GPU sharding)

    # ------------------------------------------------------------------ setup
    def setupModel(self, params_list, data, precip_mult=None, first_column=0, quiet=True):
        """``params_list``: one parameter dict per column (all on the same well,
        i.e. the same vertical grid); ``data``: the forcing table shared by the
        columns; ``precip_mult``: optional per-column precipitation multipliers."""
        if not params_list:
            raise ValueError(" EnsembleSimulation: no columns given.")
        n = len(params_list)
        mult = np.ones(n) if precip_mult is None else np.asarray(precip_mult, dtype=float)
        cols = [pack.pack_column(p, data, precip_mult=float(m), quiet=quiet)
                for p, m in zip(params_list, mult)]
        x = cols[0]["x"]
        for c in cols[1:]:
            if c["x"].shape != x.shape or not np.array_equal(c["x"], x):
                raise RuntimeError(" EnsembleSimulation: all columns must share one vertical grid.")
        m = self.mData
        m["z_grid"], m["n_cols"], m["dz"] = x, n, cols[0]["objs"]["dz"]
        m["par"] = np.stack([c["par"] for c in cols])
        m["stat"] = np.stack([c["stat"] for c in cols])
        m["forcing"] = cols[0]["forcing"]
        m["objs"] = [c["objs"] for c in cols]
        m["dim_t"] = cols[0]["forcing"]["dim_t"]
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
        self.ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])

    def _column_rng(self, local_col):
        # same stream as the reference's Simulation(name, seed): default_rng(SeedSequence(seed))
        return np.random.default_rng(np.random.SeedSequence(self.seed + self.first_column + local_col))

    def n_refresh_slots(self):
        """Tape slots a run needs: 1 + number of rows that refresh the field
        (reference simulation.py:599)."""
        m = self.mData
        f = m["forcing"]
        rows = np.arange(m["dim_t"])
        mult_max = float(m["par"][:, pack.P_PRECIP_MULT].max())
        refresh = ((mult_max * f["precipitation_cm"] > 0.5) | (rows % 48 == 0)) & (rows > 0)
        return 1 + int(refresh.sum())

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

    def initial_conditions(self, max_iter=1500):
        """Spin-up of every column on the device (reference simulation.py:389-493).
        Returns the number of spin-up iterations per column."""
        m = self.mData
        n, nx = m["n_cols"], m["z_grid"].size
        psi0 = np.stack([pack.saturated_initial_state(o) for o in m["objs"]])
        self.ctx.set_state(psi0)
        f = m["forcing"]
        nrnd0 = None
        if self.noise == "tape":
            self._rngs = [self._column_rng(c) for c in range(n)]
            nrnd0 = np.stack([r.standard_normal(nx) for r in self._rngs])
        else:
            self.ctx.set_noise_philox(self.seed, self.first_column)
        iters = self.ctx.spinup(f["zWtd_cm"][0], nrnd0, max_iter=max_iter)
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
            slots = self.n_refresh_slots()
            tape = np.empty((n, slots, nx))
            for c, r in enumerate(rngs):
                tape[c] = r.standard_normal((slots, nx))
            ctx.set_noise_tape(tape)
        else:
            ctx.set_noise_philox(self.seed, self.first_column)
        ctx.set_output(out_stride)
        ctx.set_state(m["initial_cond"])
        ctx.reset_counters()
        ctx.run(0, T)
        wscal = ctx.get_wscal()[:, :T]
        z, f = m["z_grid"], m["forcing"]
        wtd_idx = wscal[:, :, pack.W_WTD_EST].astype(np.int64)
        out = self.output
        out["wtd_est_cm"] = z[wtd_idx]
        out["abs_error"] = np.abs(f["zWtd_cm"][None, :T] - out["wtd_est_cm"])
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
            poro = np.stack([o["porosity"].profile for o in m["objs"]])
            out["S_eff"] = out["theta_vol"] / poro[:, None, :]
        out["counters"] = ctx.get_counters()
        m["final_state"] = ctx.get_state()
        return out

    # ------------------------------------------------------------- statistics
    def wtd_statistics(self, group=None):
        """Mean and standard deviation of the estimated water-table node index per
        row over ALL columns of the job.  With ``group`` (a torch.distributed
        process group, NCCL) the per-GPU moments are all-reduced over the ranks --
        the only collective of this path; it runs once, after the window loop."""
        import torch
        from .sharding import reduce_moments
        m = self.mData
        T = m["dim_t"]
        dev = torch.device("cuda", self.device)
        mom = torch.zeros(2, T, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        self.ctx.wtd_moments_dev(mom[0].data_ptr(), mom[1].data_ptr())
        self.ctx.sync()
        count = torch.full((T,), float(m["n_cols"]), dtype=torch.float64, device=dev)
        mean, std, count = reduce_moments(count, mom[0], mom[1], group)
        return mean.cpu().numpy(), std.cpu().numpy(), int(count[0].item())

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
