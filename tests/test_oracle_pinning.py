"""Pins the C restatement (oracle/siril_oracle.c) to the reference itself.

The reference ships no tests or golden vectors for the stacking path, so the
pin is (a) the reference's own stacking.c / statistics.c compiled unmodified
(oracle/_ref, built from /root/reference by oracle/Makefile) run on the same
seeded inputs, and (b) tests/golden/*.npz, outputs of oracle/_ref committed
together with the script that made them (tests/golden/make_golden.py), which
also hold on machines where /root/reference does not exist.
"""
import glob
import os

import numpy as np
import pytest

import synth

HERE = os.path.dirname(os.path.abspath(__file__))
REJ = {"none": (0, (0, 0)), "percentile": (1, (0.2, 0.1)), "sigma": (2, (4.0, 3.0)),
       "sigmedian": (3, (4.0, 3.0)), "winsorized": (4, (4.0, 3.0)), "linearfit": (5, (5.0, 2.5))}


@pytest.mark.parametrize("n,c", [(5, 1), (4, 3), (32, 1), (33, 1)])
def test_median_matches_reference(port, refrun, n, c):
    fr = synth.make_random(n, c, 12, 24, seed=n)
    r = refrun.run(fr, refrun.MEDIAN)
    p = port.stack_median(fr)
    assert r["retval"] == 0 and p["retval"] == 0
    assert np.array_equal(r["out"], p["out"])
    # D7: truncating median
    s = np.sort(fr.astype(np.int64), axis=0)
    expect = ((s[(n - 1) // 2] + s[n // 2]) // 2).astype(np.uint16)
    assert np.array_equal(p["out"], expect)


@pytest.mark.parametrize("norm", [1, 2, 3, 4])
def test_median_with_normalization_matches_reference(port, refrun, norm):
    fr = synth.make_frames(9, 1, 16, 32, seed=5)
    r = refrun.run(fr, refrun.MEDIAN, normalize=norm, want_coeff=True)
    p = port.stack_median(fr, normalize=norm, coeff=r["coeff"])
    assert np.array_equal(r["out"], p["out"])


def test_median_ser_row_order(port, refrun):
    fr = synth.make_random(6, 3, 10, 16, seed=2)
    r = refrun.run(fr, refrun.MEDIAN, row_order=1)
    p = port.stack_median(fr, row_order=1)
    assert np.array_equal(r["out"], p["out"])
    assert np.array_equal(p["out"], port.stack_median(fr)["out"][:, ::-1])


@pytest.mark.parametrize("hi", [300, 65536])
@pytest.mark.parametrize("c", [1, 3])
def test_sum_matches_reference(port, refrun, hi, c):
    fr = synth.make_random(16, c, 16, 24, seed=11, hi=hi)
    r = refrun.run(fr, refrun.SUM)
    p = port.stack_sum(fr)
    assert r["retval"] == 0
    assert np.array_equal(r["out"], p["out"])
    assert (p["out"].reshape(c, -1)[:, 0] == 0).all()      # D8: source index 0 is never read
    assert (p["maxim"] > 65535) == (hi == 65536)           # both the plain and the rescaled branch


def test_sum_with_shifts_matches_reference(port, refrun):
    fr = synth.make_random(6, 1, 20, 28, seed=4, hi=4000)
    sx = [0, 2, -3, 1, 0, -1]
    sy = [0, -2, 1, 3, -1, 0]
    for ro in (0, 1):
        r = refrun.run(fr, refrun.SUM, shiftx=sx, shifty=sy, row_order=ro)
        p = port.stack_sum(fr, shiftx=sx, shifty=sy, row_order=ro)
        assert np.array_equal(r["out"], p["out"]), ro


@pytest.mark.parametrize("take_max", [True, False])
@pytest.mark.parametrize("c", [1, 3])
def test_addmax_addmin_match_reference(port, refrun, take_max, c):
    """stack_addmax / stack_addmin (stacking.c:824-1128): image and gfit.hi, with and without
    registration shifts, both row conventions; source index 0 is never taken."""
    fr = synth.make_random(7, c, 18, 26, seed=21 + c, hi=65536)
    fr[:, :, 0, 0] = 65535 if take_max else 0               # would win if index 0 were read
    meth = refrun.ADDMAX if take_max else refrun.ADDMIN
    sx = [0, 2, -3, 1, 0, -1, 5]
    sy = [0, -2, 1, 3, -1, 0, -4]
    for ro in (0, 1):
        for shx, shy in ((None, None), (sx, sy), (sx, None)):
            r = refrun.run(fr, meth, shiftx=shx, shifty=shy, row_order=ro)
            p = port.stack_extreme(fr, take_max, shiftx=shx, shifty=shy, row_order=ro)
            assert r["retval"] == 0 and p["retval"] == 0
            assert np.array_equal(r["out"], p["out"]), (ro, shx)
            assert r["hi"] == p["hi"], (ro, shx)
    p = port.stack_extreme(fr, take_max)
    assert (p["out"].reshape(c, -1)[:, 0] == (0 if take_max else 65535)).all()   # never touched: calloc / memset 0xFF
    assert port.stack_extreme(fr[:1], take_max)["retval"] == -1                  # "select at least 2"


@pytest.mark.parametrize("rej", list(REJ))
@pytest.mark.parametrize("norm", [0, 1, 2, 3, 4])
def test_mean_matches_reference(port, refrun, rej, norm):
    """Same mini-GSL on both sides (long double sd): bit-exact image and counters.
    Pixels where the reference's "N - r <= 4" break leaves stale flags behind are history
    dependent in the reference itself (siril_oracle.c); they are excluded, and the counters
    are compared whenever no such pixel exists."""
    rj, sig = REJ[rej]
    fr = synth.make_frames(40, 1, 32, 48, seed=3, outlier_p=4e-3, zero_p=1e-3, stars=4)
    r = refrun.run(fr, refrun.MEAN, rj, norm, sig, want_coeff=True)
    p = port.stack_mean(fr, rj, sig, norm, coeff=r["coeff"])
    assert r["retval"] == 0 and p["retval"] == 0
    assert p["cap_hit"] == 0
    ok = ~p["guard_map"]
    assert ok.mean() > 0.99
    assert np.array_equal(r["out"][ok], p["out"][ok])
    if p["guard_fired"] == 0:
        assert np.array_equal(r["irej"], p["irej"])
    else:
        assert rej in ("winsorized", "linearfit")


def test_mean_rgb_and_ser_matches_reference(port, refrun):
    fr = synth.make_frames(16, 3, 16, 24, seed=9, outlier_p=4e-3)
    for ro in (0, 1):
        r = refrun.run(fr, refrun.MEAN, 2, 2, (4, 3), row_order=ro, want_coeff=True)
        p = port.stack_mean(fr, 2, (4, 3), 2, coeff=r["coeff"], row_order=ro)
        assert np.array_equal(r["out"], p["out"])
        assert np.array_equal(r["irej"], p["irej"])


def test_mean_with_shifts_matches_reference(port, refrun):
    """Registration shifts: x handled at gather (out of bounds -> literal 0), y by moving the
    read area (out of image -> zero-filled, then normalized).  One block per channel row group;
    shifts stay below the reference's block height (its own offset arithmetic is only right there)."""
    fr = synth.make_frames(12, 1, 64, 40, seed=21)
    sx = [0, 3, -2, 1, 0, -4, 2, 0, 1, -1, 0, 2]
    sy = [0, -3, 2, 4, -1, 0, 1, -2, 3, 0, -4, 1]
    for ro in (0, 1):
        for norm in (0, 3):
            r = refrun.run(fr, refrun.MEAN, 2, norm, (4, 3), shiftx=sx, shifty=sy, row_order=ro,
                           want_coeff=True, max_thread=1)
            p = port.stack_mean(fr, 2, (4, 3), norm, coeff=r["coeff"], shiftx=sx, shifty=sy, row_order=ro)
            assert p["guard_fired"] == 0
            assert np.array_equal(r["out"], p["out"]), (ro, norm)
            assert np.array_equal(r["irej"], p["irej"])


def test_statistics_match_reference(port, refrun):
    fr = synth.make_frames(4, 3, 64, 96, seed=8, zero_p=5e-3)
    for k in range(4):
        a = refrun.frame_stats(fr[k])
        b = port.frame_stats_sorted(fr[k])
        assert a == b                                       # literal restatement: bit-identical
        loc, sc, norm, _ = port.frame_stats_hist(fr[k])
        assert loc == a[0] and norm == a[2]                 # location is a bin value: exact
        assert abs(sc - a[1]) <= 1e-11 * a[1]               # count*term vs one term per sample


def test_statistics_8bit_range(port, refrun):
    """hist_size 256 when the frame's max over all layers is <= 255 (gui/histogram.c:116)."""
    fr = synth.make_random(1, 3, 32, 32, seed=1, lo=1, hi=200)
    a = refrun.frame_stats(fr[0])
    assert a[2] == 255.0
    assert port.frame_stats_sorted(fr[0]) == a
    fr[0, 2, 5, 5] = 300                                     # another layer lifts the max
    a = refrun.frame_stats(fr[0])
    assert a[2] == 65535.0
    assert port.frame_stats_sorted(fr[0]) == a
    loc, sc, norm, _ = port.frame_stats_hist(fr[0])
    assert norm == 65535.0 and loc == a[0]


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_coefficients_match_reference(port, refrun, mode):
    fr = synth.make_frames(7, 1, 48, 64, seed=13)
    r = refrun.run(fr, refrun.NORM_ONLY, normalize=mode, want_coeff=True, reference_image=2)
    st = [port.frame_stats_sorted(f) for f in fr]
    c = port.norm_coeffs(mode, [s[0] for s in st], [s[1] for s in st], ref_image=2)
    assert np.array_equal(c, r["coeff"])
    assert c[0, 2] == 0.0 and c[1, 2] == 1.0 and c[2, 2] == 1.0   # the reference frame is the identity


def test_sd_formulations_agree(port):
    """GSL's long-double sd vs the exact-integer form the GPU uses: at most +-1 ADU on a
    vanishing fraction of pixels, identical counters."""
    fr = synth.make_frames(32, 1, 64, 64, seed=17, outlier_p=4e-3)
    for rj in (2, 3, 4):
        a = port.stack_mean(fr, rj, (4, 3), sd_mode=port.SD_GSL)
        b = port.stack_mean(fr, rj, (4, 3), sd_mode=port.SD_EXACT)
        d = np.abs(a["out"].astype(np.int32) - b["out"].astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() <= 1e-4
        assert np.array_equal(a["irej"], b["irej"])


def test_golden_vectors(port):
    """Outputs of oracle/_ref committed under tests/golden (made by make_golden.py)."""
    files = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
    assert files, "no golden fixtures committed"
    for f in files:
        g = np.load(f)
        fr = g["frames"]
        kind = str(g["kind"])
        if kind == "median":
            out = port.stack_median(fr, normalize=int(g["normalize"]), coeff=g["coeff"])["out"]
        elif kind == "sum":
            out = port.stack_sum(fr)["out"]
        elif kind in ("addmax", "addmin"):
            r = port.stack_extreme(fr, kind == "addmax", shiftx=g["shiftx"], shifty=g["shifty"])
            out = r["out"]
            assert r["hi"] == int(g["hi"]), f
        else:
            r = port.stack_mean(fr, int(g["rejection"]), tuple(g["sig"]), int(g["normalize"]), coeff=g["coeff"])
            out = r["out"]
            assert np.array_equal(r["irej"], g["irej"]), f
        assert np.array_equal(out, g["out"]), f
        if "stats" in g.files:
            for k in range(fr.shape[0]):
                assert port.frame_stats_sorted(fr[k])[:2] == tuple(g["stats"][k]), f
