import os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key = "hbf10run_kernelILi256ELi2ELi2ELb1E"; fn = sys.argv[1] if len(sys.argv) > 1 else "bdf_attempt"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "hydromodel_b200/lib/libhydrob200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [k for k, l in enumerate(dis) if l.startswith("//--------------------- .text.") and key in l][0]
end = next(k for k in range(start + 1, len(dis)) if dis[k].startswith("//--------------------- "))
name = "?"; ops = collections.Counter(); calls = collections.Counter()
for l in dis[start:end]:
    m = re.match(r"^(\S+):$", l)
    if m and not m.group(1).startswith(".L_"):
        name = m.group(1).split("$")[-1]
    elif fn in name:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(\S+)", l)
        if m:
            op = m.group(1).split(".")[0]
            ops[op] += 1
            if op == "CALL":
                calls[l.split("$")[-1].strip(" ;)`")[:40]] += 1
print(sum(ops.values()), dict(ops.most_common(25)))
print(dict(calls))
