"""Build an alternative libhydrob200 for A/B measurements (HB_LIB_PATH=<out> python tools/perf_probe.py ...).
usage: python tools/build_alt.py <out.so> [-DNAME=VALUE ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hydromodel_b200 import build as hb
out = os.path.abspath(sys.argv[1])
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = [hb.nvcc_path()] + hb.NVCC_FLAGS + sys.argv[2:] + ["-o", out, hb.SRC]
res = subprocess.run(cmd, capture_output=True, text=True)
log = res.stdout + res.stderr
open(out + ".log", "w").write(" ".join(cmd) + "\n" + log)
import re
for m in re.finditer(r"Compiling entry function '(_ZN3hbf10run_kernel[^']+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log):
    print(m.group(1)[:60], "stack", m.group(2), "spill st/ld", m.group(3), m.group(4), "regs", m.group(5))
print("rc", res.returncode, out)
