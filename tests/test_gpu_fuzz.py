"""Randomised shape records (the generator of test_compile_fuzz.py: arbitrary style changes, open and self-crossing
chains, shared fills, strokes with zero and non-zero widths, morph shapes with visible round strokes) rendered on the
GPU against the oracle: edges, bin counts and pixels bit-exact."""
import numpy as np
import pytest

import corpus
from test_compile_fuzz import _random_tag

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(16))
def test_random_record_scenes_bit_exact(built_library, seed):
    rng = np.random.RandomState(9000 + seed)
    W, H = 320, 240
    sc = corpus.Scene(W, H)
    for k in range(6):
        morph = rng.rand() < 0.4
        tag = _random_tag(seed * 50 + k, morph)
        s = float(rng.uniform(0.3, 1.2))
        ang = float(rng.uniform(0, 2 * np.pi)) if rng.rand() < 0.5 else 0.0
        a, b = s * np.cos(ang), s * np.sin(ang)
        m = [float(np.float32(v)) for v in (a, a, b, -b, rng.uniform(40, W - 40) * 20.0, rng.uniform(40, H - 40) * 20.0)]
        if morph:
            sc.draw_morph(sc.add_morph(tag), m, int(rng.choice([0, 65535, rng.randint(0, 65536)])))
        else:
            sc.draw_shape(sc.add_shape(tag), m)
    ref, info = corpus.render_oracle(sc, want_debug=True)
    r, stages = corpus.make_product(sc)
    r.render(stages[0])
    out = r.get_image(premultiplied=True).data
    edges, epath = r.debug_edges(0)
    np.testing.assert_array_equal(edges, info["edges"])
    np.testing.assert_array_equal(epath, info["edge_path"])
    np.testing.assert_array_equal(r.debug_tile_counts(0), info["tile_counts"])
    r.close()
    bad = (out != ref).any(axis=2)
    assert not bad.any(), "%d px differ, first %s" % (bad.sum(), np.argwhere(bad)[:5].tolist())
