"""Top stall lines of an `ncu --page source --csv` dump: python scripts/ncu_hot.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for c in stall_cols:
    agg[c] = sum(int(r[ix[c]] or 0) for r in data)
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
top = sorted(enumerate(data), key=lambda x: -int(x[1][ix["# Samples"]] or 0))[:n]
for i, r in sorted(top):
    st = {c[6:]: int(r[ix[c]] or 0) for c in stall_cols if int(r[ix[c]] or 0)}
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {r[ix['Source']].strip()[:70]:70s} {st}")
