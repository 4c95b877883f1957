#!/usr/bin/env python
"""bench.py - throughput of the shape -> pixels hot path on the BASELINE.json workload.

Workload (config.workload): SURVEY.md 8(d) config 5 / BASELINE.json configs[4] - a synthetic stream of
10 000 random SWF shapes per frame at 1920x1080 (curves, left+right fills, solid / gradient / bitmap paints,
translucency), 64 distinct frames per step and per GPU; frames are sharded over ranks with no data-path
collective (weak scaling).  A "step" renders one batch of 64 frames.

  value     Mpixel/s of rasterized output, whole job, stages resident in HBM (swfr_batch_render), CUDA events.
  e2e       the same metric through the reference-facing C ABI call with HOST buffers: swfr_render_batch on host
            stage arrays (host flattening + H2D inside the timed region) + D2H of every finished frame to pinned
            host memory.
  roofline  dominant kernel (k_fine: per-tile coverage + paint + blend): algorithmic bytes / CUDA-event time.
  cpu_baseline  the oracle's C restatement of the reference CPU path, 1 thread, one frame of the same stream.

`--impl reference` times the reference's own CPU algorithm (oracle restatement; the TypeScript/node-canvas
original cannot run here: no node, no Rust toolchain) on all host cores, one frame of the same stream per core
and per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rasterized_Mpixel_per_s"
UNIT = "Mpixel/s"

# BASELINE.json configs[1..3]: workloads.py builds the scenes the parity tests check
EXTRA_CONFIGS = {
    "gradients1080": {"frames_per_pass": 0,
                      "workload": "60 frames at 1920x1080, one full-frame gradient shape each: linear / radial / focal x pad / "
                                  "reflect / repeat x sRGB / linear-RGB, 2-15 stops (SURVEY 8d config 2, BASELINE configs[1])"},
    "morphsweep": {"frames_per_pass": 64,
                   "workload": "flat-morph-shapes/homestuck-beta-29 at x8 (1072x720), 256 morph ratios r = 257 k in one batch "
                               "(SURVEY 8d config 3, BASELINE configs[2])"},
    "textured4k": {"frames_per_pass": 8,
                   "workload": "32 frames at 3840x2160, one full-frame bitmap-filled quad each: clipped / repeating, texel:pixel "
                               "0.25 / 1 / 2.58 / 8, corpus and 1024x1024 noise textures (SURVEY 8d config 4, BASELINE configs[3])"},
}


def hbm_peak():
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return peak, src


def kernels_sha():
    """Identity of the kernel sources this run was built from (profiles/ncu_summary.json carries the same hash for the
    build its figures were captured on)."""
    import hashlib

    h = hashlib.sha256()
    for name in ("kernels.cu", "kernels.h"):
        with open(os.path.join(ROOT, "swf_renderer_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_figures():
    """k_fine DRAM traffic and warp instructions per launch from the committed ncu capture - only when that capture
    was taken on the kernel sources this run uses."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            j = json.load(f)
        if j.get("kernels_sha") != kernels_sha():
            return None, None, "profiles/ncu_summary.json was captured on other kernel sources (%s): not quoted" % j.get("kernels_sha")
        return j.get("k_fine_dram_bytes_per_launch"), j.get("k_fine_warp_instructions_per_launch"), \
            "profiles/%s_kernels.md (ncu --set full, 16 frames per launch)" % j.get("tag")
    except Exception as e:
        return None, None, "no ncu summary (%s)" % type(e).__name__


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="distinct frames per step and per GPU")
    ap.add_argument("--shapes", type=int, default=10000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames-per-pass", type=int, default=0, help="0 = library default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--occlusion-chunks", type=int, default=0, help="0 = library default (automatic)")
    ap.add_argument("--gather", action="store_true",
                    help="N > 1 only: also time the optional gather of the ranks' finished frames onto rank 0, GPU to GPU "
                         "(NCCL over NVLink; SURVEY 8e - off the hot path, reported as its own object)")
    ap.add_argument("--quick", action="store_true", help="experiments: skip the sync-every-step leg of e2e and the extra configs")
    ap.add_argument("--config", default="stream", choices=["stream"] + list(EXTRA_CONFIGS),
                    help="workload of the line: stream = BASELINE configs[4] (the headline, with the other configs as extra "
                         "objects under `configs`); gradients1080 / morphsweep / textured4k = BASELINE configs[1] / [2] / [3] alone")
    ap.add_argument("--stroke-fraction", type=float, default=0.0,
                    help="fraction of the stream's shapes that carry a 2 px stroke (SURVEY 8d config 5, 'separate variant': 0.25)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the gradients1080 / morphsweep / textured4k objects")
    ap.add_argument("--uhd-frames", type=int, default=16,
                    help="frames of the secondary 3840x2160 measurement (same generator, radii x2); 0 = skip")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": "synthetic stream, %d random SWF shapes/frame at %dx%d (SURVEY 8d config 5, BASELINE configs[4])%s"
        % (a.shapes, a.width, a.height,
           ", %.0f %% of the shapes with a 2 px stroke" % (100 * a.stroke_fraction) if getattr(a, "stroke_fraction", 0) else ""),
        "frames_per_step_per_gpu": a.frames,
        "shapes_per_frame": a.shapes,
        "resolution": [a.width, a.height],
        "sharding": "frames over %d rank(s), no data-path collective" % n_gpus,
        "cache": "inputs larger than L2: %d distinct frames (draw lists + %.0f MB of output) per step"
        % (a.frames, a.frames * a.width * a.height * 4 / 1e6),
    }


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ----------------------------------------------------------------------------------------------------------
# CPU arms (oracle = restatement of the reference CPU path; test infrastructure used here as the baseline only)
# ----------------------------------------------------------------------------------------------------------


def _oracle_scene(frame_index, a):
    import synth
    from oracle import compile_shape as cs
    from oracle import raster

    fr = synth.SynthFrame(frame_index, a.shapes, a.width, a.height, stroke_fraction=getattr(a, "stroke_fraction", 0.0))
    b = raster._Builder({i: t for i, t in enumerate(synth.textures())})
    for i in range(fr.n):
        b.add_item(raster.add_shape_def(b, cs.compile_shape(fr.ast(i))), fr.matrix(i))
    return b, b.scene(a.width, a.height)


_WORKER = {}


def _worker_init(frame_base, a_dict):
    a = argparse.Namespace(**a_dict)
    import multiprocessing as mp

    idx = mp.current_process()._identity[0] - 1
    _WORKER["scene"] = _oracle_scene(frame_base + idx, a)


def _worker_step(_):
    from oracle import raster

    t = time.perf_counter()
    raster.render_scene(_WORKER["scene"][1])
    return time.perf_counter() - t


def cpu_baseline_single(a):
    from oracle import raster

    _, sc = _oracle_scene(0, a)
    t = time.perf_counter()
    pixels = raster.render_scene(sc)
    dt = time.perf_counter() - t
    return pixels, {
        "value": a.width * a.height / dt / 1e6,
        "unit": UNIT,
        "cores": 1,
        "kind": "port",
        "sample": "frame 0 of the same stream (1 of %d frames), C restatement of the reference CPU path, 1 thread, %.1f s"
        % (a.frames, dt),
    }


def probe_node_canvas():
    """the preferred CPU baseline is the reference's own TypeScript NodeCanvasRenderer
    (ts/src/lib/renderers/node-canvas-renderer.ts:7-24) - runnable only where `node`, the npm packages canvas /
    swf-tree / kryo and a built copy of the reference's ts/ tree exist.  Returns what was found; the C restatement
    is timed when any piece is missing (always so far: the image has no node and the GPU box no reference tree)."""
    import shutil

    found = {"node": None, "canvas": False, "swf_tree": False, "reference_ts_build": None}
    node = shutil.which("node")
    if node:
        try:
            found["node"] = subprocess.run([node, "-v"], capture_output=True, text=True, timeout=20).stdout.strip() or node
            for mod, key in (("canvas", "canvas"), ("swf-tree", "swf_tree")):
                rc = subprocess.run([node, "-e", "require(%r)" % mod], capture_output=True, timeout=30).returncode
                found[key] = rc == 0
        except Exception:
            pass
    for cand in (os.environ.get("SWFR_REFERENCE_TS", ""), "/root/reference/ts/build/lib", os.path.join(ROOT, "baseline", "_ref", "ts", "build", "lib")):
        if cand and os.path.isdir(cand):
            found["reference_ts_build"] = cand
            break
    found["usable"] = bool(found["node"] and found["canvas"] and found["swf_tree"] and found["reference_ts_build"])
    return found


def run_reference(a):
    """Reference arm: the reference's CPU algorithm on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    probe = probe_node_canvas()
    # (a usable node + canvas + reference build would be driven here, one process per core; none of the images this
    # has run in has them, so the restatement below is what gets timed - and the line says so)
    import multiprocessing as mp

    from oracle import raster

    raster.build()
    cores = max(1, len(os.sched_getaffinity(0)))
    cores = min(cores, 64)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(0, vars(a))) as pool:
        for _ in range(a.warmup):
            pool.map(_worker_step, range(cores), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_worker_step, range(cores), chunksize=1)
        dt = time.perf_counter() - t0
    px = cores * a.width * a.height * a.steps
    value = px / dt / 1e6
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": a.gpus,
        "steps": a.steps,
        "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8 (Q16 integer coverage, f32 paint)",
        "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "shapes_per_s": cores * a.shapes * a.steps / dt,
        "cpu_baseline": {
            "value": value,
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": "%d frames of the same stream per step (one per core), C restatement of the reference's "
            "TypeScript/Cairo path (node and cargo are absent, the original cannot run)" % cores,
            "node_canvas_probe": probe,
            "note": "scalar -O2 restatement without occlusion culling, several times slower than node-canvas/Cairo is likely to "
                    "be on the same cores: a reported baseline, not a claim about the reference's speed",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------


def pin_to_gpu_numa_node(local):
    """One process per GPU: keep the rank's threads and its pinned host buffers (first-touch) on the NUMA node the
    GPU's PCIe lanes hang off, so that the ranks' device->host traffic does not all cross one socket.  Best effort:
    returns a short description, or None when sysfs gives no answer."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        mine = cpus & set(os.sched_getaffinity(0))
        if not mine or mine == set(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, mine)
        return "gpu %s -> cpus %s" % (bdf, txt)
    except Exception:
        return None


def run_extra_config(name, a, local, stream, rank, world, barrier, max_over_ranks, with_cpu):
    """One of BASELINE configs[1..3] (few, large shapes - the regime SURVEY 8d expects closest to the HBM bound): the
    scene the parity test of the same name checks, all frames in one resident batch.  Returns the config's object:
    value (device-timed, stages resident), e2e (host stages in, frames out to pinned host memory), roofline of k_fine,
    cpu_baseline and parity (oracle on a sample of the frames; N = 1 only)."""
    import numpy as np
    import torch

    import workloads
    from swf_renderer_b200 import capi
    from swf_renderer_b200.renderer import _stage_arrays

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import corpus

    spec = EXTRA_CONFIGS[name]
    sc = getattr(workloads, name)()
    W, H, F = sc.width, sc.height, len(sc.frames)
    r, stages = corpus.make_product(sc, device=local, cuda_stream=stream.cuda_stream)
    if spec["frames_per_pass"]:
        r.set_option(capi.OPT_FRAMES_PER_PASS, spec["frames_per_pass"])
    arr, keep = _stage_arrays(stages)
    batch = r.create_batch((arr, keep))
    for _ in range(3):
        batch.render()
    r.sync()
    # size the timed region to ~0.5 s
    t0 = time.perf_counter()
    batch.render()
    r.sync()
    steps = int(min(2000, max(10, 0.5 / max(time.perf_counter() - t0, 1e-4))))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        batch.render()
    e1.record(stream)
    r.sync()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    stats = r.stats()
    frame_sample = r.get_image(frame=F // 2, premultiplied=True).data.copy()
    r.set_option(capi.OPT_PROFILE, 1)
    acc, passes = {}, 1
    for _ in range(3):
        batch.render()
        st = r.stage_times()
        passes = max(st["passes"], 1)
        for k, v in st["ms"].items():
            acc[k] = acc.get(k, 0.0) + v / 3
    r.set_option(capi.OPT_PROFILE, 0)
    batch.close()
    px = F * W * H
    # algorithmic bytes: segments and draw items read once, binned records written + read once, style
    # tables (4 KB ramp per gradient fill, distinct texels at most once), the frame stored once (it starts from a clear)
    n_items = stats["n_primitives"]
    style_bytes = 4096 * sum(1 for t in sc.shapes for f in t["shape"]["initial_styles"]["fill"] if f["type"].endswith("gradient"))
    tex_bytes = sum(int(b.shape[0]) * int(b.shape[1]) * 4 for b in sc.bitmaps.values())
    seg_bytes = (52 if sc.morphs else 28) * stats["n_segments"]
    alg = seg_bytes + 48 * n_items + 16 * stats["n_records"] + style_bytes + tex_bytes + 4 * px
    fpp_eff = -(-F // passes)
    fine_launches = sum(-(-min(fpp_eff, F - p * fpp_eff) // 16) for p in range(passes))  # one per slice of 16 frames
    fine_ms = acc.get("fine", 0.0) / fine_launches
    fine_bytes = (8 * stats["n_records"] + 4 * px + style_bytes + tex_bytes) / fine_launches
    peak, peak_src = hbm_peak()
    out = {
        "workload": spec["workload"],
        "resolution": [W, H],
        "frames_per_step_per_gpu": F,
        "steps": steps,
        "ms_per_step": ms,
        "value": world * px / (ms / 1e3) / 1e6,
        "unit": UNIT,
        "frames_per_s": world * F / (ms / 1e3),
        "launches_per_step": stats["kernel_launches"],
        "roofline": {
            "bound": "hbm",
            "kernel": "k_fine (per-tile coverage + paint + blend)",
            "achieved": fine_bytes / (fine_ms / 1e3) / 1e9 if fine_ms else None,
            "peak": peak,
            "peak_source": peak_src,
            "unit": "GB/s",
            "frac": fine_bytes / (fine_ms / 1e3) / 1e9 / peak if fine_ms else None,
            "traffic": None,
            "algorithmic_bytes_per_launch": fine_bytes,
            "launch_ms": fine_ms,
            "launches_per_step": fine_launches,
            "passes_per_step": passes,
            "stage_ms_per_step": acc,
            "pipeline": {"algorithmic_bytes_per_step": alg, "achieved": alg / (ms / 1e3) / 1e9, "frac": alg / (ms / 1e3) / 1e9 / peak},
        },
        "per_step": {k: stats[k] for k in ("n_primitives", "n_segments", "n_edges", "n_records", "fine_slots", "fine_records")},
    }
    # end to end: host stage arrays in, every frame out to pinned host memory, streamed over two buffers
    fb = W * H * 4
    host_out = [torch.empty(F * fb, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for i in range(2):
        r.render_stage_array(arr, F)
        r.read_frames_async(0, F, host_out[i & 1].data_ptr())
    r.sync()
    n_e2e = int(min(200, max(4, 0.5 / max(ms / 1e3 * 3, 1e-4))))
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        r.render_stage_array(arr, F)
        r.read_frames_async(0, F, host_out[i & 1].data_ptr())
    r.sync()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    out["e2e"] = {"value": world * px * n_e2e / e2e_s / 1e6, "unit": UNIT, "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e,
                  "h2d_bytes_per_step": 52 * n_items, "d2h_bytes_per_step": F * fb}
    e2e_sample = host_out[(n_e2e - 1) & 1][(F // 2) * fb:(F // 2 + 1) * fb].numpy().reshape(H, W, 4).copy()
    del host_out
    r.close()
    if with_cpu:
        # the oracle renders frames of the same scene for about ten seconds: CPU baseline, and the checker of frame F/2
        want = corpus.render_oracle(sc, frame=F // 2)
        d1 = int((frame_sample != want).any(axis=2).sum())
        d2 = int((e2e_sample != want).any(axis=2).sum())
        out["parity"] = {"frame": F // 2, "equal": d1 == 0 and d2 == 0, "px_diff": d1, "px_diff_e2e": d2,
                         "against": "oracle/raster.c, premultiplied RGBA8, bit-exact"}
        t0, n_cpu = time.perf_counter(), 0
        for f in range(F):
            corpus.render_oracle(sc, frame=f)
            n_cpu += 1
            if time.perf_counter() - t0 > 8.0:
                break
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu * W * H / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "frames 0..%d of the same scene (%d of %d), C restatement of the reference CPU path, "
                                         "1 thread, %.1f s" % (n_cpu - 1, n_cpu, F, dt)}
    return out


def run_stroked_variant(a, local, stream, rank, world, barrier, max_over_ranks, with_cpu, n_frames=16, fraction=0.25):
    """SURVEY 8d config 5, 'separate variant': the same stream with a 2 px stroke on a quarter of the shapes (butt caps,
    miter joins: outlines made at registration, composited with the per-sub-scanline coverage routine).  16 frames,
    stages resident; frame 0 is checked against the oracle at N = 1."""
    import torch

    import synth
    import swf_renderer_b200 as sw
    from swf_renderer_b200 import capi
    from swf_renderer_b200.renderer import stage_array_from_numpy, stages_from_prims

    r = sw.HeadlessRenderer(a.width, a.height, device=local, cuda_stream=stream.cuda_stream)
    r.set_option(capi.OPT_RETAIN_COMPILED, 0)
    for i, t in enumerate(synth.textures()):
        r.register_bitmap(i, t)
    prims = []
    for f in range(n_frames):
        fr = synth.SynthFrame(rank * n_frames + f, a.shapes, a.width, a.height, stroke_fraction=fraction)
        prims.append(stage_array_from_numpy(fr.register(r), fr.matrices()))
    batch = r.create_batch(stages_from_prims(prims))
    for _ in range(3):
        batch.render()
    r.sync()
    barrier()
    steps = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        batch.render()
    e1.record(stream)
    r.sync()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    st = r.stats()
    frame0 = r.get_image(frame=0, premultiplied=True).data.copy()
    batch.close()
    r.close()
    out = {"workload": "the synthetic stream with a 2 px stroke on %.0f %% of the shapes (SURVEY 8d config 5, separate variant)" % (100 * fraction),
           "resolution": [a.width, a.height], "frames_per_step_per_gpu": n_frames, "shapes_per_frame": a.shapes, "steps": steps,
           "ms_per_step": ms, "value": world * n_frames * a.width * a.height / (ms / 1e3) / 1e6, "unit": UNIT,
           "shapes_per_s": world * n_frames * a.shapes / (ms / 1e3),
           "per_step": {k: st[k] for k in ("n_primitives", "n_path_instances", "n_segments", "n_edges", "n_records", "fine_slots")}}
    if with_cpu:
        import argparse as _ap

        _, sc = _oracle_scene(0, _ap.Namespace(shapes=a.shapes, width=a.width, height=a.height, stroke_fraction=fraction))
        from oracle import raster

        want = raster.render_scene(sc)
        d = int((frame0 != want).any(axis=2).sum())
        out["parity"] = {"frame": 0, "equal": d == 0, "px_diff": d, "against": "oracle/raster.c, premultiplied RGBA8, bit-exact"}
    return out


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import synth
    import swf_renderer_b200 as sw
    from swf_renderer_b200 import capi
    from swf_renderer_b200.renderer import stage_array_from_numpy, stages_from_prims

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with_cpu = world == 1 and not a.no_cpu_baseline
    if a.config != "stream":
        # one of BASELINE configs[1..3] as the line's workload
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        obj = run_extra_config(a.config, a, local, stream, rank, world, barrier, max_over_ranks, with_cpu)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            line = {"metric": METRIC, "value": obj["value"], "unit": UNIT, "n_gpus": world, "steps": obj["steps"], "warmup": 3,
                    "ms_per_step": obj["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "u8 (Q16 integer coverage, f32 paint)", "data": "synthetic",
                    "config": {"workload": obj["workload"], "frames_per_step_per_gpu": obj["frames_per_step_per_gpu"],
                               "resolution": obj["resolution"], "sharding": "replicas of the batch over %d rank(s)" % world,
                               "cache": "%.0f MB of output per step (larger than L2)"
                               % (obj["frames_per_step_per_gpu"] * obj["resolution"][0] * obj["resolution"][1] * 4 / 1e6)},
                    "clocks": clocks, "gpu_launches": obj["launches_per_step"] * obj["steps"]}
            for k in ("e2e", "roofline", "per_step", "frames_per_s", "cpu_baseline", "parity"):
                if k in obj:
                    line[k] = obj[k]
            print(json.dumps(line), flush=True)
            if "parity" in obj and not obj["parity"]["equal"]:
                raise SystemExit("bench.py: the timed configuration differs from the oracle (see `parity` in the line)")
        if world > 1:
            dist.destroy_process_group()
        return
    r = sw.HeadlessRenderer(a.width, a.height, device=local, cuda_stream=stream.cuda_stream)
    r.set_option(capi.OPT_RETAIN_COMPILED, 0)
    if a.frames_per_pass:
        r.set_option(capi.OPT_FRAMES_PER_PASS, a.frames_per_pass)
    if a.occlusion_chunks:
        r.set_option(capi.OPT_OCCLUSION_CHUNKS, a.occlusion_chunks)
    # stage-flattening threads: the box's cores are shared by the ranks of the node
    cores = max(1, len(os.sched_getaffinity(0)))
    ranks_sharing = max(world, 1) if numa is None else max(1, (world + 1) // 2)  # ranks on this rank's cores
    host_threads = max(1, min(8, cores // ranks_sharing))
    if os.environ.get("SWFR_BENCH_HOST_THREADS"):
        host_threads = int(os.environ["SWFR_BENCH_HOST_THREADS"])
    r.set_option(capi.OPT_HOST_THREADS, host_threads)

    # ---- inputs: textures, definitions, stages (host arrays) ----
    for i, t in enumerate(synth.textures()):
        r.register_bitmap(i, t)
    prim_arrays = []
    for f in range(a.frames):
        fr = synth.SynthFrame(rank * a.frames + f, a.shapes, a.width, a.height, stroke_fraction=a.stroke_fraction)
        ids = fr.register(r)
        prim_arrays.append(stage_array_from_numpy(ids, fr.matrices()))
    stage_arr, keep = stages_from_prims(prim_arrays)
    h2d_bytes = a.frames * a.shapes * 52  # 48 B draw item + 4 B offsets per primitive
    d2h_bytes = a.frames * a.width * a.height * 4
    px_per_step = a.frames * a.width * a.height

    # ---- value: stages resident in HBM ----
    # E_tile of SURVEY 8(d) - the records the path's algorithm bins, the sum of the bit-exact tile bin counts - comes
    # from one render without occlusion culling (the chunk layout is fixed when a batch is built); the timed renders
    # skip most of those records (hidden under opaque covers).
    r.set_option(capi.OPT_OCCLUSION_CHUNKS, 1)
    full = r.create_batch((stage_arr, keep))
    full.render()
    n_records_full = r.stats()["n_records"]
    full.close()
    r.set_option(capi.OPT_OCCLUSION_CHUNKS, a.occlusion_chunks)
    batch = r.create_batch((stage_arr, keep))
    # clocks are sampled every 100 ms from the warm-up on (the same load as the timed steps) until the end of the
    # end-to-end timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(a.warmup, 3)):
        batch.render()
    r.sync()
    stats = r.stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        batch.render()
    e1.record(stream)
    r.sync()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches_per_render = r.stats()["kernel_launches"]
    launches = launches_per_render * a.steps
    # the same measurement over at least one second (clock and thermal behaviour), when K steps are shorter than that
    sustained = None
    if ms < 1000.0 and not a.quick:
        n_long = int(1000.0 / max(ms / a.steps, 1e-3)) + 1
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        for _ in range(n_long):
            batch.render()
        l1.record(stream)
        r.sync()
        barrier()
        lms = max_over_ranks(l0.elapsed_time(l1))
        sustained = {"steps": n_long, "ms_per_step": lms / n_long, "value": world * px_per_step * n_long / (lms / 1e3) / 1e6,
                     "unit": UNIT, "timed_region_s": lms / 1e3}
    # frame 0 of the batch the timed loop just rendered (compared with the oracle's frame 0 below: `parity`)
    timed_frame0 = r.get_image(frame=0, premultiplied=True).data.copy() if rank == 0 else None
    value = world * px_per_step * a.steps / (ms / 1e3) / 1e6

    # ---- roofline of the dominant kernel, CUDA events on the launching stream (separate pass) ----
    r.set_option(capi.OPT_PROFILE, 1)
    acc = {}
    n_prof = 3
    passes = 0
    for _ in range(n_prof):
        batch.render()
        st = r.stage_times()
        passes = st["passes"]
        for k, v in st["ms"].items():
            acc[k] = acc.get(k, 0.0) + v / n_prof
    r.set_option(capi.OPT_PROFILE, 0)
    stats = r.stats()
    # k_fine is launched once per slice of 16 frames of a pass (kFineSliceFrames)
    fpp_eff = -(-a.frames // max(passes, 1))
    fine_launches = sum(-(-min(fpp_eff, a.frames - p * fpp_eff) // 16) for p in range(max(passes, 1)))
    fine_ms_per_launch = acc["fine"] / max(fine_launches, 1)
    fine_bytes_per_launch = (8 * n_records_full + 4 * px_per_step) / max(fine_launches, 1)
    # B = 28 B x segments + 48 B x draw items + 2 x 8 B x E_tile + 4 x W x H per frame
    algorithmic_bytes = 28 * stats["n_segments"] + 48 * stats["n_primitives"] + 16 * n_records_full + 4 * px_per_step
    peak, peak_src = hbm_peak()
    achieved = fine_bytes_per_launch / (fine_ms_per_launch / 1e3) / 1e9
    # ncu figures are per launch of 16 frames of this stream (tools/gpu_exp.sh): scaled to this run's frames per launch
    traffic16, winst16, ncu_src = ncu_figures()
    frames_per_launch = a.frames / max(fine_launches, 1)
    traffic = traffic16 * frames_per_launch / 16.0 if traffic16 else None
    winst = winst16 * frames_per_launch / 16.0 if winst16 else None
    roofline = {
        "bound": "hbm",
        "kernel": "k_fine (per-tile coverage + paint + blend)",
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "traffic_source": ncu_src,
        "kernels_sha": kernels_sha(),
        "warp_instructions_per_launch": winst,
        # the bound the kernel actually runs against: warp instructions issued per second against 148 SMs x 4 schedulers
        # x one warp instruction per clock at the SM clock sampled during the run (filled in below, with the clocks)
        "instruction_roofline": None,
        "algorithmic_bytes_per_launch": fine_bytes_per_launch,
        "launch_ms": fine_ms_per_launch,
        "launches_per_step": fine_launches,
        "passes_per_step": passes,
        "stage_ms_per_step": acc,
        "records": {"algorithmic_E_tile_per_step": n_records_full, "binned_after_culling": stats["n_records"],
                    "read_by_k_fine": stats["fine_records"]},
        "pipeline": {
            "algorithmic_bytes_per_step": algorithmic_bytes,
            "ms_per_step": ms / a.steps,
            "achieved": algorithmic_bytes / (ms / a.steps / 1e3) / 1e9,
            "frac": algorithmic_bytes / (ms / a.steps / 1e3) / 1e9 / peak,
            "note": "whole step, passes overlapped on two streams; stage_ms_per_step is measured with the passes "
                    "serialised on one stream (SWFR_OPT_PROFILE)",
        },
    }

    # ---- e2e: host stage arrays in, finished frames out to pinned host memory ----
    # Streaming use of the C ABI: swfr_render_batch(host stages) + swfr_read_frames_async(pinned host buffer) per
    # step, two host output buffers in rotation, one swfr_sync at the end of the timed region.  Every step's stage
    # flattening, H2D copy, kernels and D2H copy of all its frames happen inside the timed region; consecutive
    # steps overlap (host flattening and PCIe traffic of one step run under the kernels of its neighbours).
    host_out = [torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for i in range(3):
        r.render_stage_array(stage_arr, a.frames)
        r.read_frames_async(0, a.frames, host_out[i & 1].data_ptr())
    r.sync()
    barrier()
    t0 = time.perf_counter()
    host_call_s = 0.0  # time the caller spends inside swfr_render_batch (stage flattening, upload and enqueue; it also
    for i in range(a.steps):  # waits there for the render before the previous one, which is long done)
        c0 = time.perf_counter()
        r.render_stage_array(stage_arr, a.frames)
        host_call_s += time.perf_counter() - c0
        r.read_frames_async(0, a.frames, host_out[i & 1].data_ptr())
    r.sync()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions (device-resident and end to end)
    e2e_value = world * px_per_step * a.steps / e2e_s / 1e6
    checksum = int(host_out[(a.steps - 1) & 1][:: 4096].to(torch.int64).sum().item())
    e2e_frame0 = host_out[(a.steps - 1) & 1][: a.width * a.height * 4].numpy().reshape(a.height, a.width, 4).copy()
    # the same call sequence without overlap between steps (sync after every step), for reference: max over ranks
    # between barriers, like every other multi-GPU number
    n_serial = 0 if a.quick else min(a.steps, 8)
    e2e_serial_ms = None
    if n_serial:
        barrier()
        t0 = time.perf_counter()
        for i in range(n_serial):
            r.render_stage_array(stage_arr, a.frames)
            r.read_frames_async(0, a.frames, host_out[i & 1].data_ptr())
            r.sync()
        e2e_serial_ms = max_over_ranks(time.perf_counter() - t0) / n_serial * 1e3
    # what the box can move: every rank copies its finished frames (the same bytes, nothing else running) from HBM to
    # its pinned host buffer, all ranks at once; device-timed between barriers, max over ranks
    d2h_ceiling = None
    if not a.quick:
        frames_dev = r.device_frames().reshape(-1)
        dst = host_out[0]
        cs = torch.cuda.Stream()
        n_copy = 6
        barrier()
        with torch.cuda.stream(cs):
            dst.copy_(frames_dev, non_blocking=True)  # warm-up
            cs.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            c0.record(cs)
            for _ in range(n_copy):
                dst.copy_(frames_dev, non_blocking=True)
            c1.record(cs)
            cs.synchronize()
        barrier()
        cms = max_over_ranks(c0.elapsed_time(c1)) / n_copy
        d2h_ceiling = {"GB/s_aggregate": world * d2h_bytes / (cms / 1e3) / 1e9, "GB/s_per_gpu": d2h_bytes / (cms / 1e3) / 1e9,
                       "ms_per_step_equivalent": cms, "value_equivalent": world * px_per_step / (cms / 1e3) / 1e6,
                       "how": "all %d rank(s) copy %d MB HBM -> pinned host concurrently, nothing else running; CUDA events, max "
                              "over ranks" % (world, d2h_bytes // 1000000)}

    # ---- optional: gather the finished frames of all ranks onto rank 0, device to device (not part of `value` / `e2e`) ----
    gather = None
    if a.gather:
        try:
            from swf_renderer_b200 import sharding

            r.sync()
            local_frames = r.device_frames()
            for _ in range(2):
                sharding.peer_gather_frames(r, dst=0)
            barrier()
            times = []
            out = None
            for _ in range(5):
                out, gms = sharding.peer_gather_frames(r, dst=0)
                if rank == 0:
                    times.append(gms)
            barrier()
            if rank == 0:
                gms = sorted(times)[len(times) // 2]
                moved = (world - 1) * local_frames.numel()
                total = world * local_frames.numel()
                gather = {"ms": gms, "frames": int(out.shape[0]), "bytes_from_peers": moved, "bytes_total": total,
                          "GB/s_from_peers": moved / (gms / 1e3) / 1e9 if world > 1 else None,
                          "GB/s_total": total / (gms / 1e3) / 1e9,
                          "how": "swfr_gather_frames: one strided cudaMemcpy2DAsync per rank (CUDA IPC peer pointers) on the "
                                 "gathering renderer's copy stream, device-timed with CUDA events; frames land in global order",
                          "first_frame_matches_rank0": bool(torch.equal(out[0], local_frames[0])),
                          "frame_of_last_rank_nonzero": bool(out[world - 1].any().item())}
            del out
            if world > 1:  # the same gather through torch.distributed (NCCL), for comparison
                n_global = world * int(local_frames.shape[0])
                sharding.gather_frames(local_frames, n_global, dst=0)
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(3):
                    o2 = sharding.gather_frames(local_frames, n_global, dst=0)
                g1.record()
                barrier()
                nms = max_over_ranks(g0.elapsed_time(g1)) / 3
                if rank == 0:
                    gather["nccl_gather_ms"] = nms
                del o2
        except Exception as e:  # the optional line must never cost the bench line
            gather = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- secondary: the same stream at 3840x2160 (BASELINE metric "at 1080p/4K"), device-resident stages ----
    uhd = None
    if a.uhd_frames > 0:
        UW, UH = 3840, 2160
        batch.close()
        ru = sw.HeadlessRenderer(UW, UH, device=local, cuda_stream=stream.cuda_stream)
        ru.set_option(capi.OPT_RETAIN_COMPILED, 0)
        ru.set_option(capi.OPT_FRAMES_PER_PASS, 8)
        for i, t in enumerate(synth.textures()):
            ru.register_bitmap(i, t)
        prims_u = []
        for f in range(a.uhd_frames):
            fr = synth.SynthFrame(rank * a.uhd_frames + f, a.shapes, UW, UH, 2.0)
            prims_u.append(stage_array_from_numpy(fr.register(ru), fr.matrices()))
        bu = ru.create_batch(stages_from_prims(prims_u))
        for _ in range(3):
            bu.render()
        ru.sync()
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_u = max(3, min(a.steps, 10))
        u0.record(stream)
        for _ in range(n_u):
            bu.render()
        u1.record(stream)
        ru.sync()
        barrier()
        ums = max_over_ranks(u0.elapsed_time(u1))
        su = ru.stats()
        uhd = {
            "resolution": [UW, UH],
            "frames_per_step_per_gpu": a.uhd_frames,
            "shapes_per_frame": a.shapes,
            "steps": n_u,
            "ms_per_step": ums / n_u,
            "value": world * a.uhd_frames * UW * UH * n_u / (ums / 1e3) / 1e6,
            "unit": UNIT,
            "shapes_per_s": world * a.uhd_frames * a.shapes * n_u / (ums / 1e3),
            "n_records": su["n_records"],
        }
        bu.close()
        ru.close()
        batch = None

    # ---- BASELINE configs[1..3] as extra objects of the line ----
    extra = {}
    if not a.quick and not a.no_extra_configs:
        if batch is not None:
            batch.close()
            batch = None
        r.close()
        r = None
        for name in EXTRA_CONFIGS:
            try:
                extra[name] = run_extra_config(name, a, local, stream, rank, world, barrier, max_over_ranks, with_cpu)
            except Exception as e:  # an extra object must never cost the headline
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        if a.stroke_fraction == 0.0:
            try:
                extra["stream_stroked25"] = run_stroked_variant(a, local, stream, rank, world, barrier, max_over_ranks, with_cpu)
            except Exception as e:
                extra["stream_stroked25"] = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0 and winst and clocks and clocks.get("sm_mhz"):
        peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
        ach = winst / (fine_ms_per_launch / 1e3)
        roofline["instruction_roofline"] = {"bound": "issue", "achieved": ach / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-inst/s",
                                            "frac": ach / peak_issue, "sm_mhz": clocks["sm_mhz"],
                                            "note": "k_fine issues on this share of all scheduler cycles; HBM is at `frac` above"}
    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": a.steps,
            "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u8 (Q16 integer coverage, f32 paint)",
            "data": "synthetic",
            "config": workload_config(a, world),
            "shapes_per_s": world * a.frames * a.shapes * a.steps / (ms / 1e3),
            "frames_per_s": world * a.frames * a.steps / (ms / 1e3),
            "clocks": clocks,
            "e2e": {
                "value": e2e_value,
                "unit": UNIT,
                "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_s / a.steps * 1e3,
                "mode": "streaming: render_batch(host stages) + read_frames_async(pinned) per step, 2 output buffers, "
                        "sync at the end",
                "host_ms_per_step_in_render_call": host_call_s / a.steps * 1e3,
                "host_threads": host_threads,
                "ms_per_step_sync_every_step": e2e_serial_ms,
                "d2h_ceiling": d2h_ceiling,
                "numa": numa,
                "checksum": checksum,
            },
            "sustained": sustained,
            "gpu_launches": launches,
            "launches_per_render": launches_per_render,
            "roofline": roofline,
            "per_step": {k: stats[k] for k in ("n_primitives", "n_segments", "n_edges", "n_slots", "n_records", "fine_slots",
                                                "fine_records", "retries")},
        }
        if uhd is not None:
            line["uhd"] = uhd
        if extra:
            line["configs"] = extra
        if gather is not None:
            line["gather"] = gather
        parity_ok = True
        if world == 1 and not a.no_cpu_baseline:
            # the oracle renders frame 0 of the same stream for the CPU baseline; its pixels check frame 0 of the timed
            # device-resident batch and frame 0 of the end-to-end host buffer, bit for bit
            want, line["cpu_baseline"] = cpu_baseline_single(a)
            d_timed = int((timed_frame0 != want).any(axis=2).sum())
            d_e2e = int((e2e_frame0 != want).any(axis=2).sum())
            parity_ok = d_timed == 0 and d_e2e == 0 and all(c.get("parity", {}).get("equal", True) for c in extra.values())
            line["parity"] = {"frame0_equal": parity_ok, "px_diff": d_timed, "px_diff_e2e": d_e2e,
                              "against": "oracle/raster.c, frame 0 of the timed batch and of the e2e host buffer, "
                                         "premultiplied RGBA8, bit-exact"}
        print(json.dumps(line), flush=True)
        if not parity_ok:
            raise SystemExit("bench.py: the timed configuration differs from the oracle (see `parity` in the line)")
    if batch is not None:
        batch.close()
    if r is not None:
        r.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
