"""Summarise an .ncu-rep: key metrics + per-function stall breakdown (SASS regions split at RET)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h} [{u}] = {v}")
for h, u, v in zip(hdr, units, vals):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.05:
        print(f"  {h.replace('smsp__average_warps_issue_stalled_','stall ').replace('_per_issue_active.ratio','')} = {float(v):.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, I, SRC = ix["# Samples"], ix["Instructions Executed"], ix["Source"]
tot = sum(int(r[S]) for r in data); toti = sum(int(r[I]) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
bounds = [0] + [k + 1 for k, r in enumerate(data) if r[SRC].strip().startswith("RET.REL.NODEC")]
print(f"total samples {tot}, warp-instructions {toti}, SASS rows {len(data)}")
for a, b in zip(bounds, bounds[1:] + [len(data)]):
    s = sum(int(r[S]) for r in data[a:b]); i = sum(int(r[I]) for r in data[a:b])
    if s < 0.003 * tot:
        continue
    st = {h: sum(int(r[ix[h]]) for r in data[a:b]) for h in stall_cols}
    top = ", ".join(f"{k[6:]} {100*v/max(s,1):.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:5])
    print(f"  SASS rows {a:5d}-{b:5d} ({b-a:4d} instr): samples {100*s/tot:5.1f}%  instr {100*i/toti:5.1f}%   [{top}]")
