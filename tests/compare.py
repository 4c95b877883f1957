"""Image comparators shared by the golden tests (SURVEY.md section 8c rules).

* compare in PREMULTIPLIED space (PNG goldens are un-premultiplied from 8-bit premultiplied storage);
* "interior" = pixels whose 3x3 golden neighbourhood is uniform in all four premultiplied channels;
* PSNR over all premultiplied channels;
* the reference's own criterion restated: pixelmatch YIQ delta, threshold 0.05, <= 0.01 % differing pixels
  (ts/src/test/node-canvas-renderer.spec.ts:182-206).
"""
import numpy as np
from scipy.ndimage import maximum_filter, minimum_filter


def premultiply_png(rgba_straight: np.ndarray) -> np.ndarray:
    g = rgba_straight.astype(np.int32)
    a = g[..., 3:4]
    pm = (g[..., :3] * a + 127) // 255
    return np.concatenate([pm, a], axis=2)


def interior_mask(gold_pm: np.ndarray) -> np.ndarray:
    m = np.ones(gold_pm.shape[:2], dtype=bool)
    for c in range(4):
        ch = gold_pm[..., c]
        m &= minimum_filter(ch, 3) == maximum_filter(ch, 3)
    return m


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean()
    return 99.0 if mse == 0 else float(10 * np.log10(255.0**2 / mse))


def stats(out_pm: np.ndarray, gold_pm: np.ndarray) -> dict:
    d = np.abs(out_pm.astype(np.int32) - gold_pm.astype(np.int32)).max(axis=2)
    inter = interior_mask(gold_pm)
    edge = ~inter
    hist = np.bincount(d[edge].ravel(), minlength=256)
    return {
        "max": int(d.max()),
        "interior_max": int(d[inter].max()) if inter.any() else 0,
        "edge_max": int(d[edge].max()) if edge.any() else 0,
        "edge_px": int(edge.sum()),
        "edge_gt2": int((d[edge] > 2).sum()),
        "edge_hist_nonzero": {int(i): int(n) for i, n in enumerate(hist) if n and i > 2},
        "psnr": psnr(out_pm, gold_pm),
    }


def pixelmatch_count(img1: np.ndarray, img2: np.ndarray, threshold: float = 0.05) -> int:
    """pixelmatch 5.1.0 core (no anti-aliasing detection => upper bound of its diff count).
    Inputs are straight RGBA8.  Pixels are blended on white before the YIQ delta, as pixelmatch does."""
    a = img1.astype(np.float64)
    b = img2.astype(np.float64)

    def blend(x):
        al = x[..., 3:4] / 255.0
        return 255.0 + (x[..., :3] - 255.0) * al

    ra, rb = blend(a), blend(b)

    def yiq(x):
        r, g, bl = x[..., 0], x[..., 1], x[..., 2]
        y = r * 0.29889531 + g * 0.58662247 + bl * 0.11448223
        i = r * 0.59597799 - g * 0.27417610 - bl * 0.32180189
        q = r * 0.21147017 - g * 0.52261711 + bl * 0.31114694
        return y, i, q

    y1, i1, q1 = yiq(ra)
    y2, i2, q2 = yiq(rb)
    delta = 0.5053 * (y1 - y2) ** 2 + 0.299 * (i1 - i2) ** 2 + 0.1957 * (q1 - q2) ** 2
    max_delta = 35215.0 * threshold * threshold
    same = (img1 == img2).all(axis=2)
    return int(((delta > max_delta) & ~same).sum())
