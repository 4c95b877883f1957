#!/usr/bin/env python
"""Summarises an `ncu --set full` report (one render pass) into profiles/: a markdown table per kernel and
profiles/ncu_summary.json (read by bench.py for roofline.traffic).

  python tools/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep gpurun_out/launches_<tag>.csv <round-tag>

Writes profiles/<round-tag>_kernels.md, profiles/<round-tag>_launches.csv (copy of the launch list) and updates
profiles/ncu_summary.json.
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * mult.get(unit, 1)


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    out = []
    for r in rows[2:]:
        def g(name):
            return num(r[col[name]]) if name in col else None

        def u(name):
            return units[col[name]] if name in col else ""

        name = r[col["Kernel Name"]].replace("swfr::<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
        name = name.split("(RenderArgs")[0].split("(const")[0].split("(unsigned")[0]
        st = sorted(((g(s) or 0.0, s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for s in stalls), reverse=True)[:3]
        rd = to_bytes(g("dram__bytes_read.sum") or 0, u("dram__bytes_read.sum"))
        wr = to_bytes(g("dram__bytes_write.sum") or 0, u("dram__bytes_write.sum"))
        dur = to_us(g("gpu__time_duration.sum") or 0, u("gpu__time_duration.sum"))
        out.append({
            "kernel": name,
            "duration_us": dur,
            "dram_read_bytes": rd,
            "dram_write_bytes": wr,
            "dram_gbs": (rd + wr) / (dur * 1e-6) / 1e9 if dur else 0,
            "dram_pct_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": g("launch__registers_per_thread"),
            "warp_instructions": g("smsp__inst_executed.sum"),
            "lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "shared_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
            "top_stalls": ["%s %.1f" % (n, v) for v, n in st],
        })
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    total = sum(k["duration_us"] for k in out) or 1
    md = ["# ncu --set full, one render pass (32 frames of the 1080p / 10 k shapes stream; k_fine per slice of 16 frames), " + tag, "",
          "Source: `%s` (cold-cache, serialised launches: compare shares, not absolutes).  "
          "lanes = average active threads per warp instruction; stalls = warps stalled per issue slot." % os.path.basename(rep), "",
          "| kernel | us | share | DRAM rd MB | DRAM wr MB | DRAM GB/s | issue % | occupancy % | regs | Minst | lanes | top stalls |",
          "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for k in out:
        md.append("| %s | %.1f | %.1f%% | %.1f | %.1f | %.0f | %.0f | %.0f | %d | %.1f | %.1f | %s |" % (
            k["kernel"], k["duration_us"], 100 * k["duration_us"] / total, k["dram_read_bytes"] / 1e6, k["dram_write_bytes"] / 1e6,
            k["dram_gbs"], k["issue_active_pct"] or 0, k["warps_active_pct"] or 0, k["registers"] or 0,
            (k["warp_instructions"] or 0) / 1e6, k["lanes_per_instruction"] or 0, ", ".join(k["top_stalls"])))
    md.append("| **total** | %.1f | | | | | | | | | | |" % total)
    with open(os.path.join(ROOT, "profiles", tag + "_kernels.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    if os.path.exists(launches):
        shutil.copy(launches, os.path.join(ROOT, "profiles", tag + "_launches.csv"))
    fine = [k for k in out if k["kernel"].startswith("k_fine")]
    import hashlib

    h = hashlib.sha256()
    for name in ("kernels.cu", "kernels.h"):
        with open(os.path.join(ROOT, "swf_renderer_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    sha = h.hexdigest()[:16]
    sha_file = os.path.splitext(rep)[0] + ".sha"  # written on the GPU box next to the capture (tools/gpu_exp.sh)
    if os.path.exists(sha_file):
        with open(sha_file) as f:
            sha = f.read().strip()
    summary = {
        "source": os.path.basename(rep),
        "tag": tag,
        # bench.py quotes the k_fine figures only when its own kernel sources hash to this value
        "kernels_sha": sha,
        "workload": "bench.py --frames 32 (one pass of the default size = one launch of every kernel, k_fine once per slice of 16 frames)",
        "k_fine_dram_bytes_per_launch": (fine[0]["dram_read_bytes"] + fine[0]["dram_write_bytes"]) if fine else None,
        "k_fine_warp_instructions_per_launch": fine[0]["warp_instructions"] if fine else None,
        "k_fine_issue_active_pct": fine[0]["issue_active_pct"] if fine else None,
        "k_fine_lanes_per_instruction": fine[0]["lanes_per_instruction"] if fine else None,
        "k_fine_duration_us_under_ncu": fine[0]["duration_us"] if fine else None,
        "kernels": [{k: v for k, v in kk.items() if k in ("kernel", "duration_us", "warp_instructions", "issue_active_pct",
                                                            "lanes_per_instruction", "dram_read_bytes", "dram_write_bytes")} for kk in out],
    }
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
