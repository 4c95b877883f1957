"""GPU parity at the BASELINE.json configurations that have no reference fixture:

* config 2 - the `triangle` and `squares` corpus geometries scaled to fill 1920x1080, every fill replaced in turn by
  linear / radial / focal(fp in {-0.75, 0, 0.5}) x spread {pad, reflect, repeat} x colour space {sRGB, linear},
  2-, 4- and 15-stop ramps, stop alpha in {255, 128}, gradient rotation {0, 30 deg};
* config 4 - the `homestuck-beta-4` quad scaled to 3840x2160 with the corpus bitmap and a seeded 1024x1024 noise +
  checker texture, {clipped, repeating} x {smoothed, unsmoothed} x texel:pixel ratio {0.25, 1, 2.58, 8};
* config 3 at throughput size - the morph shape scaled x8 (1072x720), 256 ratios in one batched launch, all frames
  bit-exact against the oracle (native size is covered by test_gpu_parity.py).

The oracle is the checker ("parity unpinned" against the reference for these paints: it has no such fixture and
throws on linear gradients, canvas-renderer.ts:332-333); GPU vs oracle is bit-exact.
"""
import numpy as np
import pytest

import corpus
import workloads

pytestmark = pytest.mark.gpu


def _render_both(sc):
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    r.close()
    return out, ref


GRAD_CASES = workloads.GRAD_CASES


@pytest.mark.parametrize("geom,kind,fp,spread,space,n_stops,alpha,rot", GRAD_CASES)
def test_config2_gradients_1080p(built_library, geom, kind, fp, spread, space, n_stops, alpha, rot):
    sc = workloads.gradient_scene((geom, kind, fp, spread, space, n_stops, alpha, rot))
    out, ref = _render_both(sc)
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:4].tolist())
    assert (out[..., 3] > 0).mean() > 0.2  # the shape really fills the frame


TEX_CASES = workloads.TEX_CASES


@pytest.mark.parametrize("which,repeating,smoothed,ratio", TEX_CASES)
def test_config4_textured_4k(built_library, which, repeating, smoothed, ratio):
    """ratio = texels per device pixel (2.58 is the corpus fixture's own minification)."""
    sc = workloads.textured_scene((which, repeating, smoothed, ratio))
    out, ref = _render_both(sc)
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:4].tolist())
    if repeating:
        assert (out[..., 3] > 0).mean() > 0.3


def test_config3_morph_sweep_x8_batched(built_library):
    """256 ratios r = 257 k in ONE batched launch at 8x the native size (1072x720): every frame bit-exact."""
    sc = workloads.morphsweep(8)
    ratios = workloads.MORPH_RATIOS
    from swf_renderer_b200 import capi

    r, stages = corpus.make_product(sc)
    r.set_option(capi.OPT_FRAMES_PER_PASS, 256)  # one set of launches for the whole sweep (and for the edge tap)
    r.render_batch(stages)
    st = r.stats()
    assert st["n_primitives"] == 256
    for f in range(256):
        out = r.get_image(frame=f, premultiplied=True).data
        if f % 8 == 0 or f == 255:  # the oracle takes ~0.1 s per frame at this size: check 33 of them pixel by pixel
            ref, info = corpus.render_oracle(sc, frame=f, want_debug=True)
            assert np.array_equal(out, ref), "ratio %d differs" % ratios[f]
            edges, _ = r.debug_edges(f)
            np.testing.assert_array_equal(edges, info["edges"])
        assert out[..., 3].max() == 255
    r.close()
