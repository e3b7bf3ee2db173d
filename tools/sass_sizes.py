"""SASS instruction count per device function of one kernel of libhydrob200.so."""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key = sys.argv[1] if len(sys.argv) > 1 else "run_kernelILi256ELi2ELi2ELb1E"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "hydromodel_b200/lib/libhydrob200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [k for k, l in enumerate(dis) if l.startswith("//--------------------- .text.") and key in l][0]
end = next(k for k in range(start + 1, len(dis)) if dis[k].startswith("//--------------------- "))
name, counts, order = "?", {}, []
for l in dis[start:end]:
    m = re.match(r"^(\S+):$", l)
    if m and not m.group(1).startswith(".L_"):
        name = m.group(1).split("$")[-1][:60]
        if name not in counts: counts[name] = 0; order.append(name)
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        counts[name] = counts.get(name, 0) + 1
tot = sum(counts.values())
for n in order:
    if counts[n]: print(f"{counts[n]:6d}  {n}")
print(f"{tot:6d}  total ({tot*16/1024:.0f} KB)")
