#!/usr/bin/env python
"""Aggregates an ncu report's source page per CUDA source line for one kernel.

  python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep k_fine [min_pct]

Prints, for every source line that executes at least min_pct % (default 0.5) of the kernel's warp instructions:
share of warp instructions, average active lanes, share of stall samples, and the source text.  Needs the
library to have been compiled with -lineinfo and the report captured with --import-source on.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    sections = {}
    agg, src = {}, {}
    hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            agg, src = sections.setdefault(r[1], ({}, {}))
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        if r[1].strip():
            src[ln] = r[1].strip()
        try:
            a = agg.setdefault(ln, [0, 0, 0])
            a[0] += int(r[ii] or 0)
            a[1] += int(r[ti] or 0)
            a[2] += int(r[si] or 0)
        except ValueError:
            pass
    for name, (agg, src) in sections.items():
        tot = sum(a[0] for a in agg.values()) or 1
        tots = sum(a[2] for a in agg.values()) or 1
        print("kernel %s: %d warp instructions, %d samples" % (name, tot, tots))
        for ln in sorted(agg):
            a = agg[ln]
            if 100.0 * a[0] / tot >= min_pct or 100.0 * a[2] / tots >= min_pct:
                print("%5d inst %5.2f%% lanes %4.1f stall %5.2f%%  %s" % (ln, 100.0 * a[0] / tot, a[1] / max(a[0], 1),
                                                                     100.0 * a[2] / tots, src.get(ln, "")[:100]))


if __name__ == "__main__":
    main()
