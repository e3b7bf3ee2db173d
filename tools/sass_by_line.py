"""Static SASS instruction count per CUDA source line range for one device function."""
import os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key = "hbf10run_kernelILi256ELi2ELi2ELb1E"; fn = sys.argv[1] if len(sys.argv) > 1 else "bdf_attempt"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "hydromodel_b200/lib/libhydrob200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [k for k, l in enumerate(dis) if l.startswith("//--------------------- .text.") and key in l][0]
end = next(k for k in range(start + 1, len(dis)) if dis[k].startswith("//--------------------- "))
name = "?"; cur = None; cnt = collections.Counter()
for l in dis[start:end]:
    m = re.match(r"^(\S+):$", l)
    if m and not m.group(1).startswith(".L_"):
        name = m.group(1).split("$")[-1]; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if fn in name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[cur] += 1
src = open(os.path.join(ROOT, "hydromodel_b200/csrc/hb_device.cuh")).read().splitlines()
tot = sum(cnt.values()); print("total", tot)
# bucket by 20-line ranges of hb_device.cuh
b = collections.Counter()
for (f, ln), v in cnt.items():
    b[(f, (ln // 25) * 25)] += v
for (f, ln), v in sorted(b.items(), key=lambda x: (x[0][0], x[0][1])):
    if v >= 40:
        text = src[ln].strip()[:70] if f == "hb_device.cuh" and ln < len(src) else ""
        print(f"{v:5d}  {f}:{ln:<5d} {text}")
