#!/usr/bin/env python
"""Prints the per-launch device times of an `ncu --metrics gpu__time_duration.sum --csv` log, with shares."""
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
out = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    out.append((r[ki].replace("swfr::<unnamed>::", "").replace("unnamed>::", "")[:70], v))
tot = sum(v for _, v in out)
for k, v in out:
    print("%-72s %9.1f us %5.1f%%" % (k, v, 100 * v / tot))
print("%-72s %9.1f us" % ("total", tot))
