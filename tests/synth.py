"""Seeded synthetic frame sequences for the parity tests (numpy, CPU).

Shapes follow SURVEY.md 8(d): per-frame background / gain / noise differences,
a smooth sky gradient, a fixed star field (some saturated), cosmic-ray style
high outliers and dead (zero) samples, so that normalization, low and high
rejection and the nullcheck exclusion are all exercised.
"""
import numpy as np


def make_frames(n, c, h, w, seed=0, outlier_p=1e-3, zero_p=1e-4, stars=40, planetary=False):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    star = np.zeros((h, w))
    srng = np.random.default_rng(seed + 12345)
    for _ in range(stars):
        sx, sy = srng.uniform(0, w), srng.uniform(0, h)
        peak = srng.uniform(2000, 90000)
        star += peak * np.exp(-((xx - sx) ** 2 + (yy - sy) ** 2) / (2 * 2.0 ** 2))
    sky = 300.0 * xx / max(w - 1, 1)
    if planetary:
        r = np.hypot(xx - w / 2, yy - h / 2) / (0.35 * min(w, h))
        disc = np.where(r < 1, 8000 + 32000 * np.sqrt(np.clip(1 - r * r, 0, 1)), 0.0)
    out = np.empty((n, c, h, w), np.uint16)
    chan_gain = [1.0] if c == 1 else [0.8, 1.0, 0.6]
    for k in range(n):
        bg = 1000 + 37 * (k % 17)
        gain = 1 + 0.02 * ((k * 7) % 11 - 5)
        sigma = 40 + (k % 5)
        for ci in range(c):
            if planetary:
                img = 200 + disc * gain * chan_gain[ci] + rng.normal(0, 200, (h, w))
            else:
                img = (bg * gain + sky + star * gain) * chan_gain[ci] + rng.normal(0, sigma, (h, w))
            hot = rng.random((h, w)) < outlier_p
            img = img + hot * rng.uniform(5000, 50000, (h, w))
            img = np.clip(np.rint(img), 0, 65535)
            img[rng.random((h, w)) < zero_p] = 0
            out[k, ci] = img.astype(np.uint16)
    return out


def make_random(n, c, h, w, seed=0, lo=0, hi=65536):
    """Uniform random samples: adversarial for sorting (ties at small ranges, 0 and 65535 present)."""
    rng = np.random.default_rng(seed)
    return rng.integers(lo, hi, size=(n, c, h, w), dtype=np.int64).astype(np.uint16)
