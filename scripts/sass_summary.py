"""Per-kernel tensor / TMA instruction counts of the built library (cuobjdump -sass; runs without a GPU):
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st (TMEM), HMMA = legacy mma.sync,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ladiff_b200", "_C", "libladiff_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", out), demangle))
OPS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "HMMA", "SYNCS", "UBLKCP")
rows, cur, arch = [], None, set()
cnt = collections.Counter()
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, dict(cnt), n_ins))
        cur, cnt, n_ins = names.get(m.group(1), m.group(1)), collections.Counter(), 0
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        n_ins += 1
        op = m.group(1).split(".")[0]
        if op in OPS:
            cnt[op] += 1
if cur:
    rows.append((cur, dict(cnt), n_ins))
print(f"# {os.path.relpath(lib, ROOT)}: arch {sorted(arch)}; {len(rows)} kernels")
print(f"{'kernel':70s} {'SASS':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for name, c, n in sorted(rows, key=lambda r: -sum(r[1].get(o, 0) for o in OPS[:5])):
    i = name.find(">(")
    short = (name[:i + 1] if i >= 0 else name.split("(")[0]).replace("void ", "").replace("(int)", "").replace("(bool)", "")
    print(f"{short[:70]:70s} {n:7d} " + " ".join(f"{c.get(o, 0):8d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL':70s} {'':7s} " + " ".join(f"{tot.get(o, 0):8d}" for o in OPS))
