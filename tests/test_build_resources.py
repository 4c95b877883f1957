"""Build-time guard on the hot kernels' resource use (cuobjdump -res-usage of the in-tree library, no GPU needed).

k_fine is instruction-bound and its speed follows occupancy: at 64 registers and no stack (no spills, no locally
indexed arrays) four 256-thread blocks are resident per SM; a change that pushes it to 80 registers (3 blocks) or
brings back a stack frame costs 15-40 % of the kernel (measurements of round 1)."""
import re
import shutil
import subprocess

import pytest

from swf_renderer_b200 import capi


def _usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-res-usage", capi.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = tuple(int(m.group(k)) for k in (2, 3, 4, 5))
    if not res:
        pytest.skip("no resource usage in cuobjdump output")
    return res


def _one(res, fragment):
    hits = {k: v for k, v in res.items() if fragment in k}
    assert hits, "kernel %s not found in the library" % fragment
    return hits


def test_library_holds_sm_100a_code_only(built_library):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-lelf", capi.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_k_fine_fits_four_blocks_per_sm_without_spills(built_library):
    res = _usage()
    # k_fine<false>: passes without stroke outlines (the benchmarked stream): 64 registers; the only stack is the pair of
    # per-warp statistics counters (swfr_stats.fine_*), touched once per composited slot - no pixel or mask ever spills
    for name, (reg, stack, shared, local) in _one(res, "6k_fineILb0").items():
        assert reg <= 64, (name, reg)
        assert stack <= 8 and local == 0, (name, stack, local)
        assert shared <= 227 * 1024 // 4
    # k_fine<true>: the only stack is the frame of the call to the out-of-line coverage routine of stroke outlines
    # (slot_coverage_sampled): registers saved around it, on that path only
    for name, (reg, stack, shared, local) in _one(res, "6k_fineILb1").items():
        assert reg <= 64, (name, reg)
        assert stack <= 32 and local == 0, (name, stack, local)
        assert shared <= 227 * 1024 // 4


def test_chunk_kernels_do_not_spill(built_library):
    res = _usage()
    for frag in ("5k_binILb0", "7k_coverE", "14k_flatten_emitILb0", "12k_path_setupE", "9k_scatterE"):
        for name, (reg, stack, shared, local) in _one(res, frag).items():
            # k_bin keeps one loop-invariant word on the stack (stored in the prologue, loaded once per 32-edge round:
            # cuobjdump -sass shows the STL / LDL outside the expansion loops); nothing else may spill
            assert stack <= (8 if "k_bin" in name else 0) and local == 0, (name, stack, local)
            assert reg <= 80, (name, reg)
