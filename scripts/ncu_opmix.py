#!/usr/bin/env python
"""Dynamic opcode mix of one kernel from an ncu report's source page.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv ; python scripts/ncu_opmix.py src.csv [n_units]
Prints executed warp-instructions and thread-instructions per opcode (optionally per unit, e.g. per pair)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
# several kernels may be concatenated: take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
w = collections.Counter(); t = collections.Counter(); samples = collections.Counter()
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or r[0] == "Address" or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name": break
        continue
    op = r[ci["Source"]].split()
    if op and op[0].startswith("@"): op = op[1:]
    name = op[0].split(".")[0] if op else "?"
    if name in ("F2F", "MUFU"): name = ".".join(op[0].split(".")[:3])
    w[name] += int(r[ci["Instructions Executed"]]); t[name] += int(r[ci["Predicated-On Thread Instructions Executed"]])
    samples[name] += int(r[ci["# Samples"]])
tw, tt, ts = sum(w.values()), sum(t.values()), sum(samples.values())
print("%-16s %12s %14s %8s %8s%s" % ("op", "warp_inst", "thread_inst", "%warp", "%samp", "  thread_inst/unit" if units else ""))
for k, v in w.most_common(40):
    print("%-16s %12d %14d %7.1f%% %7.1f%%%s" % (k, v, t[k], 100.0 * v / tw, 100.0 * samples[k] / max(ts, 1), ("  %8.2f" % (t[k] / units)) if units else ""))
print("%-16s %12d %14d%s" % ("total", tw, tt, ("  %8.2f" % (tt / units)) if units else ""))
