"""`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny workload: one JSON line with the
contract's keys, no GPU needed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--shapes", "40",
         "--width", "160", "--height", "96"],
        capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "rasterized_Mpixel_per_s" and line["unit"] == "Mpixel/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
    # the preferred baseline (node + canvas + the reference's ts/ build) is probed and reported
    probe = line["cpu_baseline"]["node_canvas_probe"]
    assert set(probe) >= {"node", "canvas", "swf_tree", "reference_ts_build", "usable"} and probe["usable"] in (True, False)


def test_committed_ncu_figures_belong_to_the_committed_kernel_sources():
    """profiles/ncu_summary.json carries the hash of the kernel sources its capture was taken on; bench.py quotes
    `roofline.traffic` and the instruction roofline only when its own sources hash to the same value - so a commit that
    touches csrc/kernels.cu or kernels.h without a fresh capture (tools/gpu_exp.sh ... full + tools/make_profiles.sh)
    fails here instead of shipping figures of another binary."""
    import json
    import os

    import bench

    with open(os.path.join(bench.ROOT, "profiles", "ncu_summary.json")) as f:
        j = json.load(f)
    assert j["kernels_sha"] == bench.kernels_sha(), "profiles/ are stale: re-capture on the current kernel sources"
    traffic, winst, src = bench.ncu_figures()
    assert traffic and winst and "r02" in src
    assert 0.3 * 259683066 < traffic < 1.2 * 259683066  # DRAM traffic of k_fine per 16-frame launch vs its algorithmic bytes
