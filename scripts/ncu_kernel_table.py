#!/usr/bin/env python
"""Per-kernel table (mean duration, DRAM MB read/written, achieved DRAM GB/s) from `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_*` logs."""
import collections
import csv
import sys

per = collections.defaultdict(lambda: collections.defaultdict(list))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    try:
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    except StopIteration:
        continue
    hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            v = float(r[ci["Metric Value"]])
        except ValueError:
            continue
        unit = r[ci["Metric Unit"]]
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        key = (r[ci["Kernel Name"]].split("(")[0].split("::")[-1][:34], r[ci["Grid Size"]])
        per[key][r[ci["Metric Name"]]].append(v * scale)
print("%-36s %-16s %5s %10s %10s %10s %9s %7s" % ("kernel", "grid", "n", "us", "rd MB", "wr MB", "GB/s", "dram%"))
for key, m in sorted(per.items()):
    avg = lambda n: sum(m[n]) / len(m[n]) if m.get(n) else 0.0
    t, rd, wr = avg("gpu__time_duration.sum"), avg("dram__bytes_read.sum"), avg("dram__bytes_write.sum")
    print("%-36s %-16s %5d %10.2f %10.3f %10.3f %9.1f %7.1f" % (key[0], key[1], len(m["gpu__time_duration.sum"]), t, rd, wr,
                                                               (rd + wr) / max(t, 1e-9) * 1e3, avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
