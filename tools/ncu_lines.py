"""Map ncu SASS-level samples to CUDA source lines via nvdisasm -g of the in-tree library.
usage: ncu_lines.py <rep> <kernel-section-substring>"""
import csv, io, re, subprocess, sys, os, tempfile, collections
rep, key = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "hydromodel_b200/lib/libhydrob200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# collect (line, inlined-at chain) for each instruction of the wanted section
start = [k for k, l in enumerate(dis) if l.startswith("//--------------------- .text.") and key in l][0]
end = next(k for k in range(start + 1, len(dis)) if dis[k].startswith("//--------------------- "))
cur = None
instrs = []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), "inlined" in m.group(3))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        instrs.append((cur, l.split("*/", 1)[1].strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, I = ix["# Samples"], ix["Instructions Executed"]
print("nvdisasm instrs", len(instrs), "ncu rows", len(data))
n = min(len(instrs), len(data))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in data)
for k in range(n):
    loc = instrs[k][0]
    key2 = (loc[0], loc[1]) if loc else ("?", 0)
    a = agg[key2]
    a[0] += int(data[k][S]); a[1] += int(data[k][I])
    for h in stall_cols:
        v = int(data[k][ix[h]])
        if v: a[2][h[6:]] += v
files = {}
def srcline(f, ln):
    if f not in files:
        for cand in (os.path.join(ROOT, "hydromodel_b200/csrc", f), ):
            if os.path.exists(cand):
                files[f] = open(cand).read().splitlines(); break
        else:
            files[f] = []
    L = files[f]
    return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
for (f, ln), (s, i, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 50]:
    top = ",".join(f"{k}:{100*v//max(s,1)}" for k, v in st.most_common(3))
    print(f"{100*s/tot:5.1f}% inst {i/1e6:8.1f}M {f}:{ln:<5d} [{top}] {srcline(f, ln)}")
