"""BASELINE.json's five workloads and a host-side (torch CPU, no CUDA, no product library)
generator of statistically identical synthetic frames.

The GPU arm of bench.py fills its frames with the library's counter-based generator
(csrc/ssc_synth.cu); the CPU reference arm must not load the product, so it draws its row band
from the same model here: per-frame background / gain / noise, sky gradient, one star per 96x96
cell (Gaussian PSF, sigma 2 px, some saturated), 1e-3 hot samples, 1e-4 dead samples; kind 1 =
limb-darkened planetary disc.  Same distribution, not the same bits.
"""
import numpy as np

SEED = 0x5349524C

# reference enums (stacking.h:14-30)
NO_REJEC, PERCENTILE, SIGMA, SIGMEDIAN, WINSORIZED, LINEARFIT = range(6)
NO_NORM, ADDITIVE, MULTIPLICATIVE, ADDITIVE_SCALING, MULTIPLICATIVE_SCALING = range(5)

CONFIGS = {
    "cfg1": dict(op="sum", n=16, c=1, h=1024, w=1024, rejection=0, sig=(0.0, 0.0), normalize=0, kind=0, row_order=0,
                 name="cfg1: sum stacking (stack_summing), 16 x 1024x1024 uint16 mono"),
    "cfg2": dict(op="median", n=32, c=1, h=4096, w=4096, rejection=0, sig=(0.0, 0.0), normalize=0, kind=0, row_order=0,
                 name="cfg2: median stack, no normalization, 32 x 4096x4096 uint16 mono"),
    "cfg3": dict(op="mean", n=64, c=3, h=4000, w=6000, rejection=SIGMA, sig=(4.0, 3.0), normalize=MULTIPLICATIVE, kind=0,
                 row_order=0,
                 name="cfg3: sigma-clip mean (4.0/3.0), multiplicative normalization, 64 x 6000x4000 uint16 RGB"),
    "cfg4": dict(op="mean", n=200, c=1, h=3520, w=4656, rejection=WINSORIZED, sig=(4.0, 3.0), normalize=ADDITIVE_SCALING,
                 kind=0, row_order=0,
                 name="cfg4: winsorized sigma-clip mean (4.0/3.0), additive+scale normalization, 200 x 4656x3520 uint16 mono"),
    "cfg5": dict(op="mean", n=1000, c=1, h=1216, w=1936, rejection=LINEARFIT, sig=(5.0, 2.5), normalize=NO_NORM, kind=1,
                 row_order=1,
                 name="cfg5: linear-fit clipping mean (5.0/2.5), no normalization, 1000 x 1936x1216 uint16 SER planetary"),
}
MODE_NAME = {"cfg1": "sum", "cfg2": "median", "cfg3": "sigma-clip mean", "cfg4": "winsorized mean", "cfg5": "linear-fit mean"}


def algorithmic_bytes(cfg, rows=None):
    """every input sample read once, every output written once, plus the second read of
    layer 0 for the normalization histograms."""
    rows = cfg["h"] if rows is None else rows
    px = rows * cfg["w"]
    stack = 2 * cfg["n"] * cfg["c"] * px + 2 * cfg["c"] * px
    stats = 2 * cfg["n"] * px if cfg["normalize"] else 0
    return stack, stats


def host_frames(cfg, row0, rows, seed=SEED, threads=None):
    """uint16 [N, C, rows, W]: storage rows [row0, row0+rows) of every frame of the workload."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    n, c, w, h, kind = cfg["n"], cfg["c"], cfg["w"], cfg["h"], cfg["kind"]
    gen = torch.Generator().manual_seed(seed ^ (row0 << 20) ^ kind)
    xx = torch.arange(w, dtype=torch.float32)[None, :]
    yy = torch.arange(row0, row0 + rows, dtype=torch.float32)[:, None]
    fixed = torch.zeros((rows, w), dtype=torch.float32)
    sky = torch.zeros((rows, w), dtype=torch.float32)
    if kind == 0:
        sky = (300.0 * xx / max(w - 1, 1)).expand(rows, w)
        srng = np.random.default_rng(seed)
        ncx, ncy = (w + 95) // 96, (h + 95) // 96
        sx = (np.arange(ncx)[None, :] + srng.random((ncy, ncx))) * 96.0
        sy = (np.arange(ncy)[:, None] + srng.random((ncy, ncx))) * 96.0
        peak = 2000.0 + 88000.0 * srng.random((ncy, ncx)) * srng.random((ncy, ncx))
        sel = (sy > row0 - 20) & (sy < row0 + rows + 20)
        for px, py, pk in zip(sx[sel], sy[sel], peak[sel]):
            x0, x1 = max(int(px) - 20, 0), min(int(px) + 21, w)
            y0, y1 = max(int(py) - 20, row0), min(int(py) + 21, row0 + rows)
            if x0 >= x1 or y0 >= y1:
                continue
            r2 = (xx[:, x0:x1] - float(px)) ** 2 + (yy[y0 - row0:y1 - row0] - float(py)) ** 2
            fixed[y0 - row0:y1 - row0, x0:x1] += float(pk) * torch.exp(-r2 * 0.125)
    else:
        rad = 0.33 * min(w, h)
        r2 = ((xx - 0.5 * w) / rad) ** 2 + ((yy - 0.5 * h) / rad) ** 2
        fixed = torch.where(r2 < 1.0, 8000.0 + 32000.0 * torch.sqrt(torch.clamp(1.0 - r2, min=0.0)), torch.zeros(()))
    out = torch.empty((n, c, rows, w), dtype=torch.int16)               # u16 payload
    chan_gain = [1.0] if c == 1 else [0.8, 1.0, 0.6]
    for k in range(n):
        for ci in range(c):
            g = torch.randn((rows, w), generator=gen, dtype=torch.float32)
            if kind == 0:
                bg = 1000.0 + 37.0 * (k % 17)
                gain = 1.0 + 0.02 * ((k * 7) % 11 - 5)
                v = (bg * gain + sky + fixed * gain) * chan_gain[ci] + (40.0 + k % 5) * g
            else:
                gain = 1.0 + 0.01 * ((k * 7) % 11 - 5)
                v = 200.0 + fixed * gain * chan_gain[ci] + 200.0 * g
            u = torch.rand((rows, w), generator=gen, dtype=torch.float32)
            v = torch.where(u < 1e-3, v + 5000.0 + 45000.0 * (u * 1e3), v)
            v = torch.clamp(torch.round(v), 0.0, 65535.0)
            v = torch.where((u > 0.5) & (u < 0.5 + 1e-4), torch.zeros(()), v)
            out[k, ci] = v.to(torch.int32).to(torch.int16)
    return out.numpy().view(np.uint16)
