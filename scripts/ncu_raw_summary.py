"""One row per kernel (first launch of each distinct name, or all with --all) out of `ncu -i file.ncu-rep --page raw --csv`:
    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x_raw.csv ; python scripts/ncu_raw_summary.py x_raw.csv [--all]
Columns: duration, DRAM bytes, tensor-pipe activity, L2 -> SM bytes, DSMEM bytes, occupancy facts."""
import csv, sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_tensor_op_umma.avg.pct_of_peak_sustained_active", "umma %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_dshared.sum", "dsmem MB"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem KB"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
ix = {h: i for i, h in enumerate(hdr)}
show_all = "--all" in sys.argv
seen = set()
print(f"{'kernel':44s} " + " ".join(f"{n:>10s}" for _, n in COLS))
for r in data:
    if len(r) < len(hdr):
        continue
    name = r[ix["Kernel Name"]].replace("void ", "")
    name = name[:name.find("(")] if "(" in name else name
    if not show_all and name in seen:
        continue
    seen.add(name)
    out = []
    for m, n in COLS:
        if m not in ix:
            out.append("-")
            continue
        try:
            v = float(r[ix[m]].replace(",", ""))
        except ValueError:
            out.append("-")
            continue
        u = units[ix[m]]
        if n.endswith("MB") or n == "us":
            v *= SCALE.get(u, 1.0)
        if n == "smem KB" and u == "byte/block":
            v /= 1024.0
        out.append(f"{v:.2f}" if abs(v) < 1000 else f"{v:.0f}")
    print(f"{name[:44]:44s} " + " ".join(f"{o:>10s}" for o in out))
