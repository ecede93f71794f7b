"""Generates tests/golden/*.npz from oracle/_ref/libsiril_ref.so -- the reference's own
stacking.c / statistics.c compiled unmodified from /root/reference (oracle/Makefile).
Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures are small on purpose (a few KB each) and are what pins the oracle and the CUDA
library on machines without the reference sources."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import synth  # noqa: E402
from oracle import port, refrun  # noqa: E402

CASES = [
    # name, kind, n, c, h, w, rejection, sig, normalize
    ("sum_n16", "sum", 16, 1, 16, 32, 0, (0, 0), 0),
    ("median_n32", "median", 32, 1, 16, 32, 0, (0, 0), 0),
    ("median_n9_addscale", "median", 9, 1, 16, 32, 0, (0, 0), 3),
    ("mean_norej_mul", "mean", 12, 1, 16, 32, 0, (0, 0), 2),
    ("mean_percentile", "mean", 16, 1, 16, 32, 1, (0.2, 0.1), 0),
    ("mean_sigma_mul_rgb", "mean", 24, 3, 16, 32, 2, (4.0, 3.0), 2),
    ("mean_sigmedian_add", "mean", 24, 1, 16, 32, 3, (4.0, 3.0), 1),
    ("mean_winsorized_addscale", "mean", 40, 1, 16, 32, 4, (4.0, 3.0), 3),
    ("mean_linearfit_mulscale", "mean", 40, 1, 16, 32, 5, (5.0, 2.5), 4),
    ("addmax_n8_rgb_shifts", "addmax", 8, 3, 16, 32, 0, (0, 0), 0),
    ("addmin_n8_shifts", "addmin", 8, 1, 16, 32, 0, (0, 0), 0),
]
SHIFTS = {  # registration shifts (regparam[reglayer][i].shiftx / .shifty) of the cases that have them
    "addmax_n8_rgb_shifts": ([0, 2, -3, 1, 0, -1, 4, -2], [0, -2, 1, 3, -1, 0, 2, -4]),
    "addmin_n8_shifts": ([0, -1, 3, 0, 2, -2, 1, 5], [0, 1, -3, 2, 0, -1, 4, -2]),
}


def main():
    only = sys.argv[1:]          # optional: names of the cases to (re)generate
    for name, kind, n, c, h, w, rej, sig, norm in CASES:
        if only and name not in only:
            continue
        seed = len(name)
        while True:
            if kind in ("sum", "addmax", "addmin"):
                fr = synth.make_random(n, c, h, w, seed=seed, hi=9000)
            else:
                fr = synth.make_frames(n, c, h, w, seed=seed, outlier_p=4e-3, zero_p=1e-3, stars=3)
            if kind != "mean":
                break
            # keep only inputs on which the reference is history independent (no "N - r <= 4" break)
            chk = port.stack_mean(fr, rej, sig, 0)
            chk2 = port.stack_mean(fr, rej, sig, norm, coeff=refrun.run(fr, refrun.NORM_ONLY, normalize=norm, want_coeff=True)["coeff"])
            if chk["guard_fired"] == 0 and chk2["guard_fired"] == 0:
                break
            seed += 100
        method = {"sum": refrun.SUM, "median": refrun.MEDIAN, "mean": refrun.MEAN,
                  "addmax": refrun.ADDMAX, "addmin": refrun.ADDMIN}[kind]
        sx, sy = SHIFTS.get(name, (None, None))
        r = refrun.run(fr, method, rej, norm, sig, want_coeff=True, shiftx=sx, shifty=sy)
        assert r["retval"] == 0, name
        stats = np.array([refrun.frame_stats(f)[:2] for f in fr]) if norm else np.zeros((0, 2))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=fr, out=r["out"], irej=r["irej"],
                            coeff=r["coeff"], kind=kind, rejection=rej, sig=np.array(sig, float),
                            normalize=norm, **({"stats": stats} if norm else {}),
                            **({"shiftx": np.array(sx, np.int32), "shifty": np.array(sy, np.int32), "hi": r["hi"]}
                               if sx is not None else {}))
        print(name, r["irej"].ravel())


if __name__ == "__main__":
    main()
