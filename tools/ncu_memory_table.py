#!/usr/bin/env python
"""Shared-memory / atomic / L2 / DRAM throughput of every kernel of an `ncu --set full` report against the device
peaks ncu itself reports (pct_of_peak_sustained_elapsed), as a markdown table.

  python tools/ncu_memory_table.py gpurun_out/prof_<tag>.ncu-rep >> profiles/<round-tag>_kernels.md
"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1.0),
    ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("DRAM rd %", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("DRAM wr %", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("smem wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("smem atom wavefronts (k)", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", 1e-3),
    ("smem atom %", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("smem atom bank conflicts (k)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", 1e-3),
    ("L2 atom sectors (k)", "lts__t_sectors_srcunit_tex_op_atom.sum", 1e-3),
    ("L2 red sectors (k)", "lts__t_sectors_srcunit_tex_op_red.sum", 1e-3),
    ("L2 atom+red %", None, 1.0),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("\n## Shared memory, atomics, L2 and DRAM against the peaks ncu reports (%s)\n" % rep.split("/")[-1])
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).split("::")[-1]
        cells = []
        for label, key, scale in COLS:
            if key is None:
                a = num(r[ix["lts__t_sectors_srcunit_tex_op_atom.sum.pct_of_peak_sustained_elapsed"]]) or 0.0
                b = num(r[ix["lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed"]]) or 0.0
                cells.append("%.2f" % (a + b))
                continue
            v = num(r[ix[key]]) if key in ix else None
            if v is None:
                cells.append("")
                continue
            if key == "gpu__time_duration.sum":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ix[key]], 1.0)
            v *= scale
            cells.append("%.1f" % v if abs(v) >= 10 else "%.2f" % v)
        print("| %s | %s |" % (name, " | ".join(cells)))


if __name__ == "__main__":
    main()
