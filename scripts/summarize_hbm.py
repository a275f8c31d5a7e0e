"""Per-kernel DRAM / L2 traffic and achieved GB/s from an ncu --csv log with gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum (and optionally lts__t_bytes.sum): python scripts/summarize_hbm.py file.csv [hbm_peak_GBs]"""
import collections, csv, json, os, re, sys

peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
rows = rows[hi:]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
acc = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}
for r in rows[1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    try:
        m, u, v = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    acc[name][m] += v * UNIT.get(u, 1.0)
    if m == "gpu__time_duration.sum":
        cnt[name] += 1
print(f"# per launch (cold caches under ncu); HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json)")
print(f"{'kernel':40s} {'launches':>8s} {'us':>8s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'GB/s':>8s} {'of peak':>8s} {'L2 MB':>8s}")
for name, d in sorted(acc.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = cnt[name]
    us = d["gpu__time_duration.sum"] / n
    rd, wr = d.get("dram__bytes_read.sum", 0.0) / n, d.get("dram__bytes_write.sum", 0.0) / n
    gbs = (rd + wr) / (us * 1e-6) / 1e9
    print(f"{name[:40]:40s} {n:8d} {us:8.2f} {rd / 1e6:11.3f} {wr / 1e6:11.3f} {gbs:8.0f} {gbs / peak:8.3f} {d.get('lts__t_bytes.sum', 0.0) / n / 1e6:8.2f}")
