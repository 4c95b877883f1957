"""GPU parity on synthetic shape streams (SURVEY.md section 8d config 5 generator): curves, left+right fills on
shared edges, rotated/scaled placement, translucent fills, all gradient kinds x spread modes x colour spaces,
clipped and repeating bitmaps - bit-exact against the oracle at sizes it finishes in seconds, and
size-independent properties at the full BASELINE size (1920x1080, 10 000 shapes)."""
import numpy as np
import pytest

import corpus
import synth

pytestmark = pytest.mark.gpu


def _scene(frame_index, n_shapes, w, h, radius_scale=1.0):
    fr = synth.SynthFrame(frame_index, n_shapes, w, h, radius_scale)
    sc = corpus.Scene(w, h)
    for i, t in enumerate(synth.textures()):
        sc.bitmaps[i] = t
    for i in range(fr.n):
        sc.draw_shape(sc.add_shape(fr.ast(i)), fr.matrix(i))
    return sc


@pytest.mark.parametrize(
    "frame_index,n_shapes,w,h,rs",
    [(0, 200, 640, 360, 0.5), (1, 400, 500, 333, 0.25), (2, 60, 320, 200, 1.0), (3, 1000, 1920, 1080, 1.0)],
)
def test_synthetic_frames_bit_exact(built_library, frame_index, n_shapes, w, h, rs):
    sc = _scene(frame_index, n_shapes, w, h, rs)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    edges, epath = r.debug_edges(0)
    np.testing.assert_array_equal(edges, info["edges"])
    np.testing.assert_array_equal(epath, info["edge_path"])
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d pixels differ, first at %s" % (bad.sum(), np.argwhere(bad)[:5].tolist())
    st = r.stats()
    assert st["n_edges"] == len(info["edges"]) and st["n_records"] == info["n_records"]
    r.close()


def test_each_fill_kind_alone(built_library):
    """One big shape per paint kind / spread / colour space so that a mismatch names its feature."""
    fr = synth.SynthFrame(7, 400, 256, 256, 0.8)
    seen = set()
    tex = synth.textures()
    for i in range(fr.n):
        key = (int(fr.fill_type[i]), int(fr.spread[i]) if 1 <= fr.fill_type[i] <= 3 else 0,
               bool(fr.linear_rgb[i]) if 1 <= fr.fill_type[i] <= 3 else False,
               bool(fr.tex_repeat[i]) if fr.fill_type[i] == 4 else False, float(fr.tex_scale[i]) if fr.fill_type[i] == 4 else 0)
        if key in seen:
            continue
        seen.add(key)
        sc = corpus.Scene(256, 256)
        for j, t in enumerate(tex):
            sc.bitmaps[j] = t
        m = fr.matrix(i)
        m[4], m[5] = 128 * 20.0, 128 * 20.0
        sc.draw_shape(sc.add_shape(fr.ast(i)), m)
        ref = corpus.render_oracle(sc)
        r, stages = corpus.make_product(sc)
        r.render(stages[0])
        out = r.get_image(premultiplied=True).data
        r.close()
        assert np.array_equal(out, ref), "paint kind %s differs in %d px" % (key, (out != ref).any(axis=2).sum())
    assert len(seen) >= 12


def test_full_size_properties(built_library):
    """1920x1080, 10 000 shapes: determinism, batch == single, tile-aligned translation equivariance."""
    import swf_renderer_b200 as sw

    W, H, N = 1920, 1080, 10000
    fr = synth.SynthFrame(0, N, W, H)
    r = sw.HeadlessRenderer(W, H)
    for j, t in enumerate(synth.textures()):
        r.register_bitmap(j, t)
    ids = fr.register(r)
    mats = fr.matrices()
    # translations on a 1/4-twip grid so that adding whole tiles is exact in float32 (Matrix2D is f32)
    mats[:, 4:6] = np.round(mats[:, 4:6] * 4) / 4

    def stage(dx_px=0.0, dy_px=0.0):
        st = sw.Stage()
        for i in range(N):
            m = mats[i].copy()
            m[4] += dx_px * 20.0
            m[5] += dy_px * 20.0
            st.display_root.append(sw.StoredShape(int(ids[i]), sw.Matrix2D(m.tolist())))
        return st

    s0 = stage()
    r.render(s0)
    a = r.get_image(premultiplied=True).data.copy()
    st = r.stats()
    # occlusion culling is on at this size: most of the geometry is hidden and never binned
    assert st["n_primitives"] == N and st["n_edges"] > N and 0 < st["fine_records"] <= st["n_records"]
    r.render(s0)
    b = r.get_image(premultiplied=True).data.copy()
    np.testing.assert_array_equal(a, b)  # deterministic despite atomics: accumulation is integer
    # the same frame inside a batch, at different frame slots
    s1 = stage(32, 16)
    r.render_batch([s1, s0, s1])
    np.testing.assert_array_equal(r.get_image(frame=1, premultiplied=True).data, a)
    c = r.get_image(frame=0, premultiplied=True).data
    np.testing.assert_array_equal(r.get_image(frame=2, premultiplied=True).data, c)
    assert a[..., 3].mean() > 200  # the stream covers nearly the whole frame
    r.close()


def test_full_size_translation_equivariance(built_library):
    """Solid fills, 1920x1080: shifting every primitive by whole tiles shifts the picture exactly (32 px and 16 px
    are exact in 24.8 fixed point and in tiles; paints that sample in float32 are excluded on purpose)."""
    import swf_renderer_b200 as sw

    W, H, N = 1920, 1080, 4000
    fr = synth.SynthFrame(5, N, W, H, solid_only=True)
    r = sw.HeadlessRenderer(W, H)
    ids = fr.register(r)
    mats = fr.matrices()
    mats[:, 4:6] = np.round(mats[:, 4:6] * 4) / 4  # 1/4-twip grid: adding whole tiles stays exact in float32
    mats[:, 0:4] = np.array([1, 1, 0, 0], dtype=np.float32)  # translate only

    def stage(dx_px, dy_px):
        st = sw.Stage()
        for i in range(N):
            m = mats[i].copy()
            m[4] += dx_px * 20.0
            m[5] += dy_px * 20.0
            st.display_root.append(sw.StoredShape(int(ids[i]), sw.Matrix2D(m.tolist())))
        return st

    r.render_batch([stage(0, 0), stage(32, 16)])
    a = r.get_image(frame=0, premultiplied=True).data
    c = r.get_image(frame=1, premultiplied=True).data
    np.testing.assert_array_equal(c[16:, 32:], a[:-16, :-32])
    r.close()


def test_streaming_renders_overlap_without_corruption(built_library):
    """swfr_render_batch / swfr_read_frames_async streamed back to back (no sync in between, two host buffers):
    every batch must come out exactly as when it is rendered alone and read synchronously."""
    import torch

    import swf_renderer_b200 as sw
    from swf_renderer_b200 import capi
    from swf_renderer_b200.renderer import stage_array_from_numpy, stages_from_prims

    W, H, N, F = 640, 360, 300, 6
    r = sw.HeadlessRenderer(W, H)
    r.set_option(capi.OPT_FRAMES_PER_PASS, 2)
    r.set_option(capi.OPT_HOST_THREADS, 3)
    for j, t in enumerate(synth.textures()):
        r.register_bitmap(j, t)
    batches = []
    for k in range(3):
        prims = []
        for f in range(F):
            fr = synth.SynthFrame(100 + k * F + f, N, W, H, 0.5)
            prims.append(stage_array_from_numpy(fr.register(r), fr.matrices()))
        batches.append(stages_from_prims(prims))
    # reference: one batch at a time, synchronous reads
    want = []
    for arr, keep in batches:
        r.render_stage_array(arr, F)
        want.append(np.stack([r.get_image(frame=f, premultiplied=True).data.copy() for f in range(F)]))
    assert not np.array_equal(want[0], want[1])
    # streamed: 5 renders in flight back to back over 2 output buffers
    fb = W * H * 4
    out = [torch.zeros(F * fb, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    order = [0, 1, 2, 1, 0]
    for i, k in enumerate(order):
        r.render_stage_array(batches[k][0], F)
        r.read_frames_async(0, F, out[i & 1].data_ptr())
    r.sync()
    np.testing.assert_array_equal(out[0].numpy().reshape(F, H, W, 4), want[order[4]])
    np.testing.assert_array_equal(out[1].numpy().reshape(F, H, W, 4), want[order[3]])
    # partial range + sync per step still works
    r.render_stage_array(batches[2][0], F)
    r.read_frames_async(1, 3, out[0].data_ptr())
    r.sync()
    np.testing.assert_array_equal(out[0].numpy()[: 3 * fb].reshape(3, H, W, 4), want[2][1:4])
    r.close()


def test_working_memory_growth_and_rerun(built_library):
    """SWFR_OPT_DEBUG_TINY_ARENA: every working array (edges, slots, records, lists, rows, staging) starts far too
    small, so the render overflows stage after stage, the host grows the arena and re-runs the pass - the result must
    be the oracle's, also for a batch whose frames were being copied out while a pass was re-run."""
    import torch

    import swf_renderer_b200 as sw
    from swf_renderer_b200 import capi
    from swf_renderer_b200.renderer import stage_array_from_numpy, stages_from_prims

    sc = _scene(11, 300, 640, 360, 0.5)
    ref = corpus.render_oracle(sc)
    r, stages = corpus.make_product(sc)
    r.set_option(capi.OPT_DEBUG_TINY_ARENA, 1)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    st = r.stats()
    assert st["retries"] >= 2, st
    np.testing.assert_array_equal(out, ref)
    r.render(stages[0])  # second time: the arena is large enough now
    assert r.stats()["retries"] == 0
    np.testing.assert_array_equal(r.get_image(premultiplied=True).data, ref)
    r.close()

    # streamed batch with async read-back while passes overflow and are re-run
    W, H, N, F = 320, 200, 150, 4
    r = sw.HeadlessRenderer(W, H)
    for j, t in enumerate(synth.textures()):
        r.register_bitmap(j, t)
    prims, want = [], []
    for f in range(F):
        fr = synth.SynthFrame(300 + f, N, W, H, 0.5)
        prims.append(stage_array_from_numpy(fr.register(r), fr.matrices()))
    arr, keep = stages_from_prims(prims)
    r.set_option(capi.OPT_FRAMES_PER_PASS, 2)
    r.render_stage_array(arr, F)
    want = np.stack([r.get_image(frame=f, premultiplied=True).data.copy() for f in range(F)])
    r.set_option(capi.OPT_DEBUG_TINY_ARENA, 1)
    buf = torch.zeros(F * W * H * 4, dtype=torch.uint8, pin_memory=True)
    r.render_stage_array(arr, F)
    r.read_frames_async(0, F, buf.data_ptr())
    r.sync()
    assert r.stats()["retries"] >= 2
    np.testing.assert_array_equal(buf.numpy().reshape(F, H, W, 4), want)
    r.close()


@pytest.mark.parametrize("chunks", [2, 4, 8])
def test_occlusion_chunks_do_not_change_pixels(built_library, chunks):
    """Depth-chunk occlusion culling (SWFR_OPT_OCCLUSION_CHUNKS): geometry under an opaque full-tile cover found by
    an upper chunk is neither flattened nor binned - the pixels must stay the oracle's, the binned records shrink,
    and the debug taps (which need complete lists) refuse to run."""
    import swf_renderer_b200 as sw
    from swf_renderer_b200 import capi
    from swf_renderer_b200.renderer import SwfrError

    sc = _scene(21, 700, 640, 360, 1.0)  # large shapes: deep overdraw
    ref = corpus.render_oracle(sc)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    full = r.stats()
    np.testing.assert_array_equal(r.get_image(premultiplied=True).data, ref)
    r.set_option(capi.OPT_OCCLUSION_CHUNKS, chunks)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    st = r.stats()
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:5].tolist())
    assert st["n_records"] < full["n_records"] * 0.8, (st["n_records"], full["n_records"])
    assert st["fine_records"] == full["fine_records"] and st["fine_slots"] == full["fine_slots"]
    with pytest.raises(SwfrError):
        r.debug_edges(0)
    # a batch: frames cull independently (covers of one frame must not leak into another)
    sc2 = _scene(22, 500, 640, 360, 1.0)
    ref2 = corpus.render_oracle(sc2)
    r2, st2 = corpus.make_product(sc2)
    r2.close()
    sid = {id(t): r.register_shape(t) for t in sc2.shapes}
    stage2 = sw.Stage([sw.StoredShape(sid[id(sc2.shapes[idx])], sw.Matrix2D(m)) for _, idx, m, _, _ in sc2.frames[0]])
    empty = sw.Stage([])
    r.render_batch([stages[0], stage2, empty, stages[0]])
    np.testing.assert_array_equal(r.get_image(frame=0, premultiplied=True).data, ref)
    np.testing.assert_array_equal(r.get_image(frame=1, premultiplied=True).data, ref2)
    assert not r.get_image(frame=2, premultiplied=True).data.any()
    np.testing.assert_array_equal(r.get_image(frame=3, premultiplied=True).data, ref)
    # growth / re-run with culling on
    r.set_option(capi.OPT_DEBUG_TINY_ARENA, 1)
    r.render_batch([stage2, stages[0]])
    assert r.stats()["retries"] >= 1
    np.testing.assert_array_equal(r.get_image(frame=0, premultiplied=True).data, ref2)
    np.testing.assert_array_equal(r.get_image(frame=1, premultiplied=True).data, ref)
    r.close()


def _random_placement_scene(seed, n_shapes, w, h):
    """Shapes of the synthetic stream under hostile placement: scales from 1/200 to 400, any rotation, shear,
    mirroring, centres up to several frames away from the viewport - most geometry is clipped, some edges are
    clamped at +-32768 px, many shapes are sub-pixel."""
    rng = np.random.RandomState(seed)
    fr = synth.SynthFrame(1000 + seed, n_shapes, w, h, 0.5)
    sc = corpus.Scene(w, h)
    for i, t in enumerate(synth.textures()):
        sc.bitmaps[i] = t
    for i in range(fr.n):
        s = float(np.exp(rng.uniform(np.log(0.005), np.log(400.0))))
        if rng.rand() < 0.5:
            s = float(rng.uniform(0.3, 3.0))
        ang = rng.uniform(0, 2 * np.pi)
        sx, sy = s * rng.uniform(0.2, 1.0), s * rng.choice([-1.0, 1.0])
        shear = rng.uniform(-0.5, 0.5)
        a, b = sx * np.cos(ang), sx * np.sin(ang)
        c, d = -sy * np.sin(ang) + shear * a, sy * np.cos(ang) + shear * b
        tx = rng.uniform(-2.0, 3.0) * w * 20.0 if rng.rand() < 0.3 else rng.uniform(0, w) * 20.0
        ty = rng.uniform(-2.0, 3.0) * h * 20.0 if rng.rand() < 0.3 else rng.uniform(0, h) * 20.0
        # Matrix2D order: scale_x, scale_y, rotate_skew0, rotate_skew1, tx, ty  (x' = m0 x + m3 y + m4, y' = m2 x + m1 y + m5)
        m = [float(np.float32(v)) for v in (a, d, b, c, tx, ty)]
        sc.draw_shape(sc.add_shape(fr.ast(i)), m)
    return sc


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_hostile_placements_bit_exact(built_library, seed):
    """Extreme scales, shear, mirroring, far off-screen and clamped geometry: edges, bin counts and pixels still equal
    the oracle's, with and without occlusion chunks."""
    from swf_renderer_b200 import capi

    sc = _random_placement_scene(seed, 160, 400, 300)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    edges, epath = r.debug_edges(0)
    np.testing.assert_array_equal(edges, info["edges"])
    np.testing.assert_array_equal(epath, info["edge_path"])
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:5].tolist())
    r.set_option(capi.OPT_OCCLUSION_CHUNKS, 3)
    r.render(stages[0])
    np.testing.assert_array_equal(r.get_image(premultiplied=True).data, ref)
    r.close()


@pytest.mark.parametrize("frame_index,n_shapes,w,h,rs", [(31, 3000, 1920, 1080, 1.0), (32, 1500, 3840, 2160, 2.0)])
def test_automatic_culling_at_full_resolution_matches_oracle(built_library, frame_index, n_shapes, w, h, rs):
    """Frames of 1024 or more items switch occlusion culling on by themselves (4 depth chunks, unordered edges, only
    visible geometry flattened): the pixels at 1080p and at 4K must still be the oracle's."""
    from swf_renderer_b200.renderer import SwfrError

    sc = _scene(frame_index, n_shapes, w, h, rs)
    ref = corpus.render_oracle(sc)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    st = r.stats()
    with pytest.raises(SwfrError):  # culling was on: the taps refuse
        r.debug_tile_counts(0)
    r.close()
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:5].tolist())
    assert st["fine_records"] <= st["n_records"] and st["n_edges"] > 0


@pytest.mark.parametrize("w,h", [(1, 1), (17, 3), (333, 77), (15, 16), (16, 15), (130, 1)])
def test_odd_viewport_sizes(built_library, w, h):
    """Viewports that are not multiples of the 16 px tile or of the 4-pixel store width (partial tiles, scalar
    stores), down to a single pixel."""
    from swf_renderer_b200 import capi

    sc = _scene(40 + w, 60, w, h, 0.3)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    assert out.shape == ref.shape == (h, w, 4)
    np.testing.assert_array_equal(out, ref)
    r.set_option(capi.OPT_OCCLUSION_CHUNKS, 2)
    r.render_batch([stages[0], stages[0], stages[0]])
    for f in range(3):
        np.testing.assert_array_equal(r.get_image(frame=f, premultiplied=True).data, ref)
    straight = r.get_image(frame=1).data  # un-premultiplied read-back with a row stride of 4 * w
    assert straight.shape == (h, w, 4) and (straight[..., 3] == ref[..., 3]).all()
    r.close()


def _oracle_frame(frame_index, n_shapes, w, h, rs=1.0):
    """The oracle's picture of one frame of the synthetic stream (the construction bench.py's cpu_baseline uses)."""
    from oracle import compile_shape as cs
    from oracle import raster

    fr = synth.SynthFrame(frame_index, n_shapes, w, h, rs)
    b = raster._Builder({i: t for i, t in enumerate(synth.textures())})
    for i in range(fr.n):
        b.add_item(raster.add_shape_def(b, cs.compile_shape(fr.ast(i))), fr.matrix(i))
    return raster.render_scene(b.scene(w, h))


@pytest.mark.parametrize(
    "w,h,rs,n_frames,fpp,check",
    [
        # bench.py's headline configuration: 10 000 shapes per frame at 1080p, ONE batch, library defaults (4 depth
        # chunks, 32 frames per pass, passes alternating over two arenas / streams); one frame of passes 0, 1 and 2
        (1920, 1080, 1.0, 65, 0, (0, 40, 64)),
        # bench.py's 4K line: radii x2, 8 frames per pass; one frame of passes 0 and 1
        (3840, 2160, 2.0, 9, 8, (3, 8)),
    ],
)
def test_benchmarked_configuration_matches_oracle(built_library, w, h, rs, n_frames, fpp, check):
    """The configuration bench.py times (BASELINE configs[4]), pixel for pixel against the oracle: multi-pass batch of
    10 000-shape frames through create_batch / render (device-resident stages, `value`) AND through
    render_stage_array + read_frames_async into pinned host memory (`e2e`)."""
    import torch

    import swf_renderer_b200 as sw
    from swf_renderer_b200 import capi
    from swf_renderer_b200.renderer import stage_array_from_numpy, stages_from_prims

    N = 10000
    r = sw.HeadlessRenderer(w, h)
    r.set_option(capi.OPT_RETAIN_COMPILED, 0)
    if fpp:
        r.set_option(capi.OPT_FRAMES_PER_PASS, fpp)
    for j, t in enumerate(synth.textures()):
        r.register_bitmap(j, t)
    prims = []
    for f in range(n_frames):
        fr = synth.SynthFrame(f, N, w, h, rs)
        prims.append(stage_array_from_numpy(fr.register(r), fr.matrices()))
    arr, keep = stages_from_prims(prims)
    batch = r.create_batch((arr, keep))
    batch.render()
    r.sync()  # the first render sizes the working memory (it may overflow and be re-run)
    batch.render()
    batch.render()  # two renders in flight back to back, as in bench.py's timed loop
    st = r.stats()
    assert st["retries"] == 0 and st["n_primitives"] == N * n_frames
    want = {f: _oracle_frame(f, N, w, h, rs) for f in check}
    for f in check:
        out = r.get_image(frame=f, premultiplied=True).data
        bad = (out != want[f]).any(axis=2)
        assert not bad.any(), "resident batch, frame %d: %d px differ, first %s" % (f, bad.sum(), np.argwhere(bad)[:5].tolist())
    batch.close()
    # the end-to-end path: host stage arrays in, frames out to pinned host memory, two renders in flight
    fb = w * h * 4
    host = [torch.zeros(n_frames * fb, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for i in range(2):
        r.render_stage_array(arr, n_frames)
        r.read_frames_async(0, n_frames, host[i].data_ptr())
    r.sync()
    for i in range(2):
        frames = host[i].numpy().reshape(n_frames, h, w, 4)
        for f in check:
            assert np.array_equal(frames[f], want[f]), "streamed render %d, frame %d differs from the oracle" % (i, f)
    r.close()


def test_slot_memory_scales_with_area_times_count(built_library):
    """1000 translucent full-screen paths at 3840x2160: slot memory is one slot per (path, tile of its bbox) - 32.4 M
    slots, far beyond the start-up heuristics.  The render must come out right after at most a few grow-and-re-run
    rounds (every overflowing counter reports the size it needs), not thrash the retry limit."""
    import swf_renderer_b200 as sw

    W, H, N = 3840, 2160, 1000
    big = 20 * 4200
    tag = {
        "id": 1,
        "bounds": {"x_min": -200, "x_max": big, "y_min": -200, "y_max": big},
        "shape": {
            "initial_styles": {"fill": [{"type": "solid", "color": {"r": 200, "g": 40, "b": 90, "a": 64}}], "line": []},
            "records": [
                {"type": "style-change", "move_to": {"x": -200, "y": -200}, "right_fill": 1},
                {"type": "edge", "delta": {"x": big + 200, "y": 0}},
                {"type": "edge", "delta": {"x": 0, "y": big + 200}},
                {"type": "edge", "delta": {"x": -big - 200, "y": 0}},
                {"type": "edge", "delta": {"x": 0, "y": -big - 200}},
            ],
        },
    }
    r = sw.HeadlessRenderer(W, H)
    sid = r.register_shape(tag)
    st = sw.Stage([sw.StoredShape(sid, sw.Matrix2D([1.0, 1.0, 0.0, 0.0, 0.0, 0.0])) for _ in range(N)])
    r.render(st)
    out = r.get_image(premultiplied=True).data
    stats = r.stats()
    assert stats["retries"] <= 3, stats
    assert stats["n_slots"] == N * ((W + 15) // 16) * ((H + 15) // 16)

    def mul_un8(a, b):
        t = a * b + 0x80
        return ((t >> 8) + t) >> 8

    # the solid colour goes through Cairo's 16-bit premultiplied path, then N times source-over
    af = float(np.float32(64 / 255.0))
    a16 = int(af * 65535.0 + 0.5)
    src = [int(((c / 255.0) * af) * 65535.0 + 0.5) >> 8 for c in (200, 40, 90)] + [a16 >> 8]
    px = [0, 0, 0, 0]
    for _ in range(N):
        px = [s + mul_un8(d, 255 - src[3]) for s, d in zip(src, px)]
    assert (out == np.array(px, dtype=np.uint8)).all(), (out[0, 0].tolist(), px)
    r.render(st)  # working memory is large enough now
    assert r.stats()["retries"] == 0
    r.close()


@pytest.mark.parametrize("n_shapes,w,h,rs", [(300, 640, 360, 0.5), (1500, 1920, 1080, 1.0)])
def test_stroked_stream_variant_bit_exact(built_library, n_shapes, w, h, rs):
    """SURVEY 8d config 5, separate variant: a 2 px stroke on a quarter of the shapes.  Static strokes become outlines
    at registration (butt caps, miter joins) and are composited with the non-zero rule per sub-scanline (stroke outlines
    overlap themselves); frames of 1024 or more items also go through occlusion culling."""
    fr = synth.SynthFrame(50 + n_shapes, n_shapes, w, h, rs, stroke_fraction=0.25)
    sc = corpus.Scene(w, h)
    for i, t in enumerate(synth.textures()):
        sc.bitmaps[i] = t
    for i in range(fr.n):
        sc.draw_shape(sc.add_shape(fr.ast(i)), fr.matrix(i))
    assert fr.stroked.sum() > n_shapes // 8
    ref = corpus.render_oracle(sc)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    st = r.stats()
    assert st["n_path_instances"] > n_shapes  # the strokes are paths of their own
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:5].tolist())
    # the fast registration path of the benchmark builds the same definitions
    r2 = type(r)(w, h)
    for j, t in enumerate(synth.textures()):
        r2.register_bitmap(j, t)
    from swf_renderer_b200.renderer import stage_array_from_numpy, stages_from_prims

    arr, keep = stages_from_prims([stage_array_from_numpy(fr.register(r2), fr.matrices())])
    r2.render_stage_array(arr, 1)
    assert np.array_equal(r2.get_image(premultiplied=True).data, ref)
    r.close()
    r2.close()
