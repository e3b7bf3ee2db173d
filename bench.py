#!/usr/bin/env python3
"""
bench.py -- throughput of the Richards time-stepping hot path.

One "step" = one pass of the window loop of Simulation.run over one ensemble of
columns (BASELINE.json configs[1]: 1,024 columns per GPU perturbing n / a0 /
psi_sat around the default input_parameters.json, shared 365-row synthetic daily
forcing, nx = 401, host noise tape drawn in the reference's RNG order, shared
spun-up initial state).  Metric: simulated column-years per second
(1 column-year = 365 windows), whole job over all GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference ...                          the CPU restatement
                                                                of the reference's path
For N > 1 launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hydromodel_b200 import pack                                   # noqa: E402
from hydromodel_b200.synthetic import (synthetic_forcing, default_params,   # noqa: E402
                                       ensemble_member)

SITE = os.path.join(ROOT, "hydromodel_b200", "model_parameters", "site_information.json")
ROWS_PER_YEAR = 365
F_CELL = 470.0          # canonical FP64 flop per midpoint closure+flux


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# dram__bytes_read.sum (352.7 MB) + dram__bytes_write.sum (269.5 MB) of run_kernel<256,2,2,true> for
# (columns, rows, out_stride, nx) below: 1.5x the 413 MB of algorithmic bytes -- the time-sliced
# scheduler hands state and noise field of a column over through HBM every 16 rows; 1.0 GB/s in
# total, 1.6e-4 of the HBM peak.  (Before the shared-memory layout was cut to 18 vectors the same
# launch wrote 686 MB: spill-stack lines evicted from a smaller L1.)
NCU_TRAFFIC = {"workload": (1024, 365, 30, 401), "bytes": 352691712 + 269521152,
               "source": "profiles/ncu_r01_bench_run_kernel_summary.txt"}


# ----------------------------------------------------------------------------- workload
def build_workload(n_cols, first_col, n_rows, total_cols):
    """Pack `n_cols` ensemble members (global ids first_col..) and the shared forcing."""
    forcing = synthetic_forcing(n_rows)
    base = default_params(SITE)
    pars, stats = [], []
    col = None
    for c in range(n_cols):
        p, mult = ensemble_member(base, first_col + c, max(total_cols, 1024))
        col = pack.pack_column(p, forcing, precip_mult=mult)
        pars.append(col["par"]); stats.append(col["stat"])
    default_col = pack.pack_column(base, forcing)
    return {"x": col["x"], "par": np.stack(pars), "stat": np.stack(stats),
            "forcing": col["forcing"], "default": default_col}


def noise_tape(n_cols, first_col, forcing, nx, seed=0):
    """Standard normals in the reference's draw order for Simulation(seed=c)
    started from an IC file: run draw, then one draw per refreshing row."""
    rows = np.arange(forcing["dim_t"])
    refresh = ((forcing["precipitation_cm"] > 0.5) | (rows % 48 == 0)) & (rows > 0)
    slots = 1 + int(refresh.sum())
    tape = np.empty((n_cols, slots, nx))
    for c in range(n_cols):
        rng = np.random.default_rng(np.random.SeedSequence(seed + first_col + c))
        tape[c] = rng.standard_normal((slots, nx))
    return tape


def algorithmic_flops(counters, nx):
    """N_rhs*nx*F_cell + N_newton*14*nx + N_lu*24*nx + N_step*40*nx."""
    c = counters.sum(axis=0).astype(float)
    return (c[pack.C_NRHS] * nx * F_CELL + c[pack.C_NNEWTON] * 14 * nx +
            c[pack.C_NLU] * 24 * nx + (c[pack.C_NSTEPS] + c[pack.C_NREJECT]) * 40 * nx)


def algorithmic_bytes(n_cols, n_rows, nx, n_out, slots):
    """per column-window 32 B scalars + 8*nx per noise refresh +
    (32*nx + 32) per output row; plus the once-per-column state / statics."""
    return n_cols * (32.0 * n_rows + 8.0 * nx * slots + (32.0 * nx + 32.0) * n_out +
                     8.0 * nx * (2 + pack.NSTATIC))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:                     # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows),
                "reasons": reasons}


# ----------------------------------------------------------------------------- CPU arm
def _cpu_column(args):
    import oracle
    x, grp, par, stat, f, tape, ic, n_rows = args
    t0 = time.perf_counter()
    psi, wscal, _, stats, _ = oracle.run(x, grp, par, stat, f["precipitation_cm"][:n_rows],
                                         f["wtd_idx"][:n_rows], f["daylight"][:n_rows],
                                         f["atm"][:n_rows], tape, ic, out_stride=30)
    return time.perf_counter() - t0, stats


def cpu_baseline(wl, ic, n_rows, n_sample_cols, threads):
    """The CPU restatement of the reference's path (oracle/hb_oracle.c: the same
    algorithm in C, one column per host thread) on a bounded sample."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    oracle.lib()
    grp, _ = pack.fd_groups(wl["x"].size)
    tape = noise_tape(n_sample_cols, 0, wl["forcing"], wl["x"].size)
    jobs = [(wl["x"], grp, wl["par"][c % wl["par"].shape[0]], wl["stat"][c % wl["par"].shape[0]],
             wl["forcing"], tape[c], ic, n_rows) for c in range(n_sample_cols)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:       # ctypes releases the GIL
        res = list(ex.map(_cpu_column, jobs))
    wall = time.perf_counter() - t0
    windows = n_sample_cols * (n_rows - 1)
    stats = np.sum([r[1] for r in res], axis=0)
    return {"value": windows / ROWS_PER_YEAR / wall, "unit": "column-years/s", "cores": threads,
            "kind": "port",
            "sample": f"{n_sample_cols} columns x {n_rows - 1} windows of the same workload, "
                      f"{threads} host threads, oracle/hb_oracle.c (C restatement of the reference's "
                      f"SciPy-BDF path; the reference's own Python runs ~18x slower per core, BASELINE.md)",
            "windows_per_s": windows / wall, "rhs_evals_per_s": float(stats[pack.C_NRHS]) / wall,
            "wall_s": wall}


def default_ic(ctx_factory, wl):
    """Shared initial state: device spin-up of the default-parameter column
    (Simulation.initial_conditions), cached per process.  Untimed setup."""
    d = wl["default"]
    from hydromodel_b200 import capi
    ctx = capi.Context(d["x"].size, 1, ctx_factory)
    grp, ng = pack.fd_groups(d["x"].size)
    ctx.set_grid(d["x"], grp, ng)
    ctx.set_columns(d["par"][None], d["stat"][None])
    f = d["forcing"]
    ctx.set_forcing(f["precipitation_cm"], f["wtd_idx"], f["daylight"], f["atm"])
    ctx.set_state(pack.saturated_initial_state(d["objs"])[None])
    nr = np.random.default_rng(np.random.SeedSequence(1234)).standard_normal(d["x"].size)
    iters = ctx.spinup(f["zWtd_cm"][0], nr[None], max_iter=1500)
    ic = ctx.get_state()[0]
    ctx.close()
    return ic, int(iters[0])


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--columns", type=int, default=1024, help="columns per GPU")
    ap.add_argument("--rows", type=int, default=365, help="forcing rows (365 = 1 column-year)")
    ap.add_argument("--out-stride", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(a.gpus, world)
    config = {"workload": "configs[1]: %d-column ensemble per GPU perturbing n/a0/psi_sat "
                          "(SeedSequence(1234) members), default input_parameters.json, Well 10 "
                          "(nx=401), vrettas_fung + Stratified, %d synthetic daily rows, shared "
                          "spun-up IC, host noise tape" % (a.columns, a.rows),
              "columns_per_gpu": a.columns, "rows": a.rows, "nx": 401,
              "rows_per_column_year": ROWS_PER_YEAR, "out_stride": a.out_stride,
              "l2": "256 MiB buffer written between timed steps (L2 flush)",
              "rtol": 1e-3, "atol": 1e-3}

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        wl = build_workload(min(a.columns, 64), 0, a.rows, a.columns * n_gpus)
        ic = np.load(os.path.join(ROOT, "hydromodel_b200", "model_parameters", "ic_well10_default.npy"))
        n_sample = max(4 * threads, 16)
        for _ in range(max(a.warmup, 1)):
            cpu_baseline(wl, ic, min(a.rows, 8), threads, threads)
        vals = [cpu_baseline(wl, ic, min(a.rows, 121), n_sample, threads) for _ in range(a.steps)]
        best = max(vals, key=lambda v: v["value"])
        mean = float(np.mean([v["value"] for v in vals]))
        best = dict(best, value=mean)
        print(json.dumps({
            "impl": "reference", "metric": "column_years_per_s", "value": mean,
            "unit": "column-years/s", "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * float(np.mean([v["wall_s"] for v in vals])),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "cpu_baseline": best,
            "e2e": {"value": mean, "unit": "column-years/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ CUDA arm
    import torch
    from hydromodel_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback "
                         "(use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    first_col = rank * a.columns
    t0 = time.time()
    wl = build_workload(a.columns, first_col, a.rows, a.columns * n_gpus)
    nx = wl["x"].size
    f = wl["forcing"]
    tape = noise_tape(a.columns, first_col, f, nx)
    ic_path = os.path.join(ROOT, "hydromodel_b200", "model_parameters", "ic_well10_default.npy")
    ic, spin_iters = default_ic(local, wl)
    if rank == 0:
        log(f"[bench] setup {time.time() - t0:.1f}s; default-column spin-up {spin_iters} iterations")
        try:
            ref_ic = np.load(ic_path)
            log("[bench] max |IC - shipped IC| = %.3e" % np.abs(ic - ref_ic).max())
        except OSError:
            pass

    # pinned host staging buffers (e2e copies come from pinned memory)
    def pinned(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        return t, t.numpy()
    keep = []
    host = {}
    for name, arr in (("par", wl["par"]), ("stat", wl["stat"]), ("tape", tape),
                      ("ic", np.tile(ic, (a.columns, 1))), ("precip", f["precipitation_cm"]),
                      ("atm", f["atm"])):
        t, v = pinned(arr); keep.append(t); host[name] = v
    grp, ng = pack.fd_groups(nx)
    ctx = capi.Context(nx, a.columns, local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_grid(wl["x"], grp, ng)

    def upload():
        ctx.set_columns(host["par"], host["stat"])
        ctx.set_forcing(host["precip"], f["wtd_idx"], f["daylight"], host["atm"])
        ctx.set_noise_tape(host["tape"])
        ctx.set_output(a.out_stride)
        ctx.set_state(host["ic"])

    def reset_state():
        ctx.set_noise_tape(host["tape"])       # the run consumes / damps the tape state
        ctx.set_state(host["ic"])
        ctx.reset_counters()

    upload()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak_fp64 = ctx.measure_fp64_peak()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only timing: inputs resident in HBM -------------------------------
    for _ in range(a.warmup):
        reset_state(); flush.fill_(1)
        with torch.cuda.stream(stream):
            ctx.run(0, a.rows)
        stream.synchronize()
    sampler = ClockSampler(local); sampler.start()
    launches0 = ctx.launch_count
    kernel_ms = []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        reset_state(); flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            ctx.run(0, a.rows)
            e1.record(stream)
        e1.synchronize()
        kernel_ms.append(e0.elapsed_time(e1))
    barrier()
    wall_kernel_loop = time.perf_counter() - wall0
    launches = ctx.launch_count - launches0
    counters = ctx.get_counters()
    clocks = sampler.summary()
    step_ms = float(np.mean(kernel_ms))
    log(f"[bench] rank {rank}: kernel ms per step {['%.1f' % v for v in kernel_ms]}, clocks {clocks}")
    if dist is not None:
        t = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
    windows = a.columns * (a.rows - 1)
    value = n_gpus * windows / ROWS_PER_YEAR / (step_ms * 1e-3)

    # ---- end to end through the C-ABI with host buffers ---------------------------
    n_out = ctx.n_output_rows
    h2d = sum(host[k].nbytes for k in ("par", "stat", "tape", "ic", "precip", "atm")) + \
        2 * 4 * a.rows
    d2h = a.columns * (a.rows * pack.NWSCAL * 8 + n_out * pack.NFIELD * nx * 8 + nx * 8)
    e2e_s = []
    res = {}                                   # pinned host buffers the results are read into
    for name, shape in (("wscal", (a.columns, a.rows, pack.NWSCAL)),
                        ("fields", (a.columns, n_out, pack.NFIELD, nx)), ("state", (a.columns, nx))):
        t, v = pinned(np.empty(shape)); keep.append(t); res[name] = v
    for it in range(2 + a.steps):
        barrier()
        t0 = time.perf_counter()
        upload()
        with torch.cuda.stream(stream):
            ctx.run(0, a.rows)
        wscal = ctx.get_wscal(out=res["wscal"])
        fields = ctx.get_fields(out=res["fields"])
        psi_end = ctx.get_state(out=res["state"])
        torch.cuda.synchronize()
        if it >= 2:
            e2e_s.append(time.perf_counter() - t0)
    e2e_step = float(np.mean(e2e_s))
    if dist is not None:
        t = torch.tensor([e2e_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_step = float(t.item())
    e2e_value = n_gpus * windows / ROWS_PER_YEAR / e2e_step

    # ---- ensemble statistics: the one collective of this path (outside the timing) --
    mom = torch.zeros(2, a.rows, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.wtd_moments_dev(mom[0].data_ptr(), mom[1].data_ptr())
    ctx.sync()
    if dist is not None:
        dist.all_reduce(mom)
    wtd_mean_last = float(mom[0, -1].item()) / (a.columns * n_gpus)

    # ---- roofline ------------------------------------------------------------------
    flops = algorithmic_flops(counters, nx)
    slots = tape.shape[1]
    abytes = algorithmic_bytes(a.columns, a.rows, nx, n_out, slots)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ach_tf = flops / (step_ms * 1e-3) / 1e12
    # DRAM bytes of one run_kernel launch from the committed `ncu --set full` capture of this
    # very workload (dram__bytes_read.sum + dram__bytes_write.sum); other workloads: not captured
    traffic, traffic_src = None, None
    if (a.columns, a.rows, a.out_stride, nx) == NCU_TRAFFIC["workload"]:
        traffic, traffic_src = NCU_TRAFFIC["bytes"], NCU_TRAFFIC["source"]
    ach_gbs = abytes / (step_ms * 1e-3) / 1e9
    csum = counters.sum(axis=0).astype(float)
    result = {
        "metric": "column_years_per_s", "value": value, "unit": "column-years/s",
        "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config,
        "windows_per_s": n_gpus * windows / (step_ms * 1e-3),
        "rhs_evals_per_s": n_gpus * csum[pack.C_NRHS] / (step_ms * 1e-3),
        "per_window": {"nfev": csum[pack.C_NFEV] / csum[pack.C_NWINDOWS],
                       "njev": csum[pack.C_NJEV] / csum[pack.C_NWINDOWS],
                       "nlu": csum[pack.C_NLU] / csum[pack.C_NWINDOWS],
                       "steps": csum[pack.C_NSTEPS] / csum[pack.C_NWINDOWS],
                       "rhs": csum[pack.C_NRHS] / csum[pack.C_NWINDOWS],
                       "failed_attempts": csum[pack.C_NFAIL]},
        "e2e": {"value": e2e_value, "unit": "column-years/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_step},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": peak_fp64, "unit": "TFLOP/s",
                     "frac": ach_tf / peak_fp64, "traffic": traffic,
                     "traffic_unit": "DRAM bytes per launch (read + write)", "traffic_source": traffic_src,
                     "peak_source": "hb_measure_fp64_peak: register-resident DFMA loop measured in "
                                    "this run (MEASURED_PEAKS.json has no FP64 figure)",
                     "flops_per_launch": flops,
                     "note": "algorithmic FP64 flop of SURVEY 8(d) from the kernel's own counters; "
                             "the path is an SMEM-resident FP64 solver, neither HBM- nor tensor-bound"},
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach_gbs / hbm_peak, "bytes_per_launch": abytes,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
        "ensemble": {"wtd_node_mean_last_row": wtd_mean_last,
                     "collective": "all_reduce(sum) of WTD moments, NCCL" if dist is not None else "none (1 GPU)"},
    }
    if rank == 0 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = cpu_baseline(wl, ic, min(a.rows, 121), max(4 * threads, 16), threads)
        result["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(result))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
