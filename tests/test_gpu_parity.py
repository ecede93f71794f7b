"""Parity of the CUDA library with the oracle -- the parity tests proper (`-m gpu`).

Everything goes through the C ABI (ssc_stack_* / ssc_dev_* in libsiril_stack_cuda.so).
Bars (BASELINE.json north_star): sum and median bit-exact; the floating-point rejection modes
bit-exact against the oracle evaluated with the same sigma formulation, and against the
reference formulation (GSL's x87 long double) at most +-1 ADU on <= 0.01 % of the pixels with
identical rejection counters.
"""
import glob
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

REJ = {"none": (0, (0, 0)), "percentile": (1, (0.2, 0.1)), "sigma": (2, (4.0, 3.0)),
       "sigmedian": (3, (4.0, 3.0)), "winsorized": (4, (4.0, 3.0)), "linearfit": (5, (5.0, 2.5))}


def close_enough(a, b, frac=1e-4):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max() <= 1 and (d != 0).sum() <= max(1, int(frac * d.size))


# --------------------------------------------------------------------------- median
@pytest.mark.parametrize("n,c,h,w", [(2, 1, 8, 8), (3, 1, 9, 17), (5, 3, 12, 24), (32, 1, 40, 64),
                                     (33, 1, 16, 40), (64, 1, 16, 48), (100, 1, 8, 24), (257, 1, 4, 16)])
def test_median_bit_exact(gpu, port, n, c, h, w):
    fr = synth.make_random(n, c, h, w, seed=n * 7 + c)
    res = gpu.stack_median(gpu.Sequence(fr))
    assert np.array_equal(res.data, port.stack_median(fr)["out"])


def test_median_ties_and_extremes(gpu, port):
    fr = synth.make_random(16, 1, 16, 32, seed=1, lo=0, hi=3)          # heavy ties
    fr[:, 0, 0, :8] = 65535
    fr[:8, 0, 1, :8] = 0
    fr[8:, 0, 1, :8] = 65535                                            # (0 + 65535) / 2 truncated
    res = gpu.stack_median(gpu.Sequence(fr))
    assert np.array_equal(res.data, port.stack_median(fr)["out"])
    assert res.data[0, 1, 0] == 32767 and res.data[0, 0, 0] == 65535


@pytest.mark.parametrize("norm", [1, 2, 3, 4])
def test_median_normalized(gpu, port, norm):
    fr = synth.make_frames(12, 1, 48, 64, seed=5)
    res = gpu.stack_median(gpu.Sequence(fr), normalize=norm, want_coeff=True)
    assert np.array_equal(res.data, port.stack_median(fr, normalize=norm, coeff=res.coeff)["out"])


def test_median_ser_order_and_subset(gpu, port):
    fr = synth.make_random(9, 3, 10, 16, seed=2)
    res = gpu.stack_median(gpu.Sequence(fr, row_order=gpu.ROWS_SER))
    assert np.array_equal(res.data, port.stack_median(fr, row_order=1)["out"])
    idx = [7, 1, 4, 2, 8]                                               # image_indices: a filtered, reordered list
    res = gpu.stack_median(gpu.Sequence(fr), image_indices=idx)
    assert np.array_equal(res.data, port.stack_median(fr[idx])["out"])


# --------------------------------------------------------------------------- sum
@pytest.mark.parametrize("hi", [300, 65536])
@pytest.mark.parametrize("c,w", [(1, 32), (3, 24), (1, 21)])
def test_sum_bit_exact(gpu, port, hi, c, w):
    fr = synth.make_random(16, c, 16, w, seed=11, hi=hi)
    res = gpu.stack_summing(gpu.Sequence(fr))
    p = port.stack_sum(fr)
    assert np.array_equal(res.data, p["out"])
    assert (res.data.reshape(c, -1)[:, 0] == 0).all()
    assert res.hi == min(p["maxim"], 65535)


def test_sum_shifts_and_ser(gpu, port):
    fr = synth.make_random(6, 1, 20, 32, seed=4, hi=4000)
    sx = [0, 2, -3, 1, 0, -1]
    sy = [0, -2, 1, 3, -1, 0]
    for ro in (0, 1):
        res = gpu.stack_summing(gpu.Sequence(fr, shiftx=sx, shifty=sy, row_order=ro))
        assert np.array_equal(res.data, port.stack_sum(fr, shiftx=sx, shifty=sy, row_order=ro)["out"]), ro
        res = gpu.stack_summing(gpu.Sequence(fr, row_order=ro))
        assert np.array_equal(res.data, port.stack_sum(fr, row_order=ro)["out"]), ro


def test_sum_filter_callback(gpu, port):
    fr = synth.make_random(8, 1, 8, 16, seed=6, hi=500)
    keep = [0, 2, 3, 6]
    crit = gpu.FILTER_FN(lambda seq, i, p: int(i in keep))
    res = gpu.stack_summing(gpu.Sequence(fr), filtering_criterion=crit)
    assert np.array_equal(res.data, port.stack_sum(fr[keep])["out"])


# --------------------------------------------------------------------------- addmax / addmin
@pytest.mark.parametrize("take_max", [True, False])
@pytest.mark.parametrize("c,w", [(1, 32), (3, 24), (1, 21)])
def test_addmax_addmin_bit_exact(gpu, port, take_max, c, w):
    """stack_addmax / stack_addmin (stacking.c:824-1128): vector path (w % 8 == 0) and general path."""
    fr = synth.make_random(9, c, 16, w, seed=31, hi=65536)
    fr[:, :, 0, 0] = 65535 if take_max else 0               # source index 0 is never taken
    op = gpu.stack_addmax if take_max else gpu.stack_addmin
    res = op(gpu.Sequence(fr))
    p = port.stack_extreme(fr, take_max)
    assert np.array_equal(res.data, p["out"])
    assert (res.data.reshape(c, -1)[:, 0] == (0 if take_max else 65535)).all()
    assert res.hi == p["hi"]
    sx = [0, 2, -3, 1, 0, -1, 5, 0, -20]
    sy = [0, -2, 1, 3, -1, 0, -4, 17, 0]                     # one frame shifted out of the image entirely
    for ro in (0, 1):
        res = op(gpu.Sequence(fr, shiftx=sx, shifty=sy, row_order=ro))
        p = port.stack_extreme(fr, take_max, shiftx=sx, shifty=sy, row_order=ro)
        assert np.array_equal(res.data, p["out"]), ro
        assert res.hi == p["hi"]
        res = op(gpu.Sequence(fr, row_order=ro))
        assert np.array_equal(res.data, port.stack_extreme(fr, take_max, row_order=ro)["out"]), ro


def test_addmax_filter_and_errors(gpu, port):
    fr = synth.make_random(8, 1, 8, 16, seed=6, hi=500)
    keep = [0, 2, 3, 6]
    crit = gpu.FILTER_FN(lambda seq, i, p: int(i in keep))
    res = gpu.stack_addmax(gpu.Sequence(fr), filtering_criterion=crit)
    assert np.array_equal(res.data, port.stack_extreme(fr[keep], True)["out"])
    with pytest.raises(gpu.StackError) as e:
        gpu.stack_addmin(gpu.Sequence(fr), image_indices=[0])
    assert e.value.code == -1                                # "No frame selected for stacking (select at least 2)"


# --------------------------------------------------------------------------- mean with rejection
@pytest.mark.parametrize("rej", list(REJ))
@pytest.mark.parametrize("norm", [0, 1, 2, 3, 4])
def test_mean_all_modes(gpu, port, rej, norm):
    rj, sig = REJ[rej]
    fr = synth.make_frames(40, 1, 32, 48, seed=3, outlier_p=4e-3, zero_p=1e-3, stars=4)
    res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), rj, sig, norm)
    coeff = res.coeff
    same = port.stack_mean(fr, rj, sig, norm, coeff=coeff, sd_mode=port.SD_EXACT)
    assert np.array_equal(res.data, same["out"])                        # same formulation: bit-exact
    assert np.array_equal(res.irej, same["irej"])
    assert res.guard_fired == same["guard_fired"] and res.cap_hit == 0
    gsl = port.stack_mean(fr, rj, sig, norm, coeff=coeff, sd_mode=port.SD_GSL)
    assert close_enough(res.data, gsl["out"])                           # reference formulation
    if same["guard_fired"] == 0:
        assert np.array_equal(res.irej, gsl["irej"])


@pytest.mark.parametrize("rej", ["sigma", "winsorized", "linearfit", "percentile"])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 17, 64, 130])
def test_mean_frame_counts(gpu, port, rej, n):
    rj, sig = REJ[rej]
    fr = synth.make_frames(n, 1, 16, 32, seed=n, outlier_p=1e-2, zero_p=2e-3, stars=2)
    res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), rj, sig)
    same = port.stack_mean(fr, rj, sig, sd_mode=port.SD_EXACT)
    assert np.array_equal(res.data, same["out"])
    assert np.array_equal(res.irej, same["irej"])
    assert res.guard_fired == same["guard_fired"]                      # small N: the guard does fire


def test_mean_rgb_ser_and_reference_image(gpu, port):
    fr = synth.make_frames(16, 3, 16, 24, seed=9, outlier_p=4e-3)
    for ro in (0, 1):
        res = gpu.stack_mean_with_rejection(gpu.Sequence(fr, row_order=ro, reference_image=3), 2, (4, 3), 2)
        same = port.stack_mean(fr, 2, (4, 3), 2, coeff=res.coeff, row_order=ro, sd_mode=port.SD_EXACT)
        assert np.array_equal(res.data, same["out"])
        assert np.array_equal(res.irej, same["irej"])
        assert res.coeff[1, 3] == 1.0                                   # the reference frame is the identity


def test_mean_with_shifts(gpu, port):
    fr = synth.make_frames(12, 1, 64, 40, seed=21)
    sx = [0, 3, -2, 1, 0, -4, 2, 0, 1, -1, 0, 2]
    sy = [0, -3, 2, 4, -1, 0, 1, -2, 3, 0, -4, 1]
    for ro in (0, 1):
        for norm in (0, 3):
            res = gpu.stack_mean_with_rejection(gpu.Sequence(fr, shiftx=sx, shifty=sy, row_order=ro), 2, (4, 3), norm)
            same = port.stack_mean(fr, 2, (4, 3), norm, coeff=res.coeff, shiftx=sx, shifty=sy, row_order=ro,
                                   sd_mode=port.SD_EXACT)
            assert np.array_equal(res.data, same["out"]), (ro, norm)
            assert np.array_equal(res.irej, same["irej"])


def test_mean_extremes(gpu, port):
    """Saturated, all-zero and two-valued columns; negative sigma parameters (everything is flagged)."""
    fr = synth.make_random(20, 1, 8, 32, seed=5)
    fr[:, 0, 0] = 65535
    fr[:, 0, 1] = 0
    fr[:10, 0, 2] = 0
    fr[10:, 0, 2] = 65535
    fr[:, 0, 3] = 1
    for rej in REJ:
        rj, sig = REJ[rej]
        for s in (sig, (-1.0, -1.0)):
            res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), rj, s)
            same = port.stack_mean(fr, rj, s, sd_mode=port.SD_EXACT)
            assert np.array_equal(res.data, same["out"]), (rej, s)
            assert np.array_equal(res.irej, same["irej"]), (rej, s)
            assert res.guard_fired == same["guard_fired"]


# --------------------------------------------------------------------------- normalization statistics
@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_coefficients(gpu, port, mode):
    fr = synth.make_frames(9, 3, 96, 128, seed=13, zero_p=3e-3)
    c = gpu.compute_normalization(gpu.Sequence(fr, reference_image=2), mode)
    hist_stats = [port.frame_stats_hist(f) for f in fr]
    expect = port.norm_coeffs(mode, [s[0] for s in hist_stats], [s[1] for s in hist_stats], ref_image=2)
    assert np.array_equal(c, expect)                                    # same algorithm, same summation order
    sorted_stats = [port.frame_stats_sorted(f) for f in fr]
    ref = port.norm_coeffs(mode, [s[0] for s in sorted_stats], [s[1] for s in sorted_stats], ref_image=2)
    assert np.allclose(c, ref, rtol=1e-11, atol=1e-9)                   # the reference's sort-based form
    if mode in (1, 2):
        assert np.array_equal(c, ref)                                   # location only: exact


def test_coefficients_8bit_range_and_zeros(gpu, port):
    fr = synth.make_random(4, 3, 32, 32, seed=1, lo=0, hi=200)          # max <= 255: hist_size 256
    c = gpu.compute_normalization(gpu.Sequence(fr), 3)
    st = [port.frame_stats_sorted(f) for f in fr]
    assert all(s[2] == 255.0 for s in st)
    assert np.allclose(c, port.norm_coeffs(3, [s[0] for s in st], [s[1] for s in st]), rtol=1e-11)
    fr[1, 2, 3, 3] = 300                                                # another layer lifts one frame's max
    c = gpu.compute_normalization(gpu.Sequence(fr), 3)
    st = [port.frame_stats_sorted(f) for f in fr]
    assert st[1][2] == 65535.0 and st[0][2] == 255.0
    assert np.allclose(c, port.norm_coeffs(3, [s[0] for s in st], [s[1] for s in st]), rtol=1e-11)


def test_all_zero_frame_is_an_error(gpu):
    fr = synth.make_frames(4, 1, 16, 16, seed=2)
    fr[2] = 0
    with pytest.raises(gpu.StackError) as e:
        gpu.stack_mean_with_rejection(gpu.Sequence(fr), 2, (4, 3), 1, want_coeff=False)
    assert e.value.code == -1                                           # the stackers: -1 (stacking.c:1337-1340)
    with pytest.raises(gpu.StackError) as e:
        gpu.compute_normalization(gpu.Sequence(fr), 1)
    assert e.value.code == 1                                            # compute_normalization itself: 1 (:161,189)


# --------------------------------------------------------------------------- golden fixtures (made from oracle/_ref)
def test_golden_fixtures(gpu):
    files = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
    assert files
    for f in files:
        g = np.load(f)
        fr, kind = g["frames"], str(g["kind"])
        seq = gpu.Sequence(fr)
        if kind == "median":
            res = gpu.stack_median(seq, normalize=int(g["normalize"]))
            assert np.array_equal(res.data, g["out"]), f
        elif kind == "sum":
            assert np.array_equal(gpu.stack_summing(seq).data, g["out"]), f
        elif kind in ("addmax", "addmin"):
            seq = gpu.Sequence(fr, shiftx=g["shiftx"], shifty=g["shifty"])
            res = (gpu.stack_addmax if kind == "addmax" else gpu.stack_addmin)(seq)
            assert np.array_equal(res.data, g["out"]), f
            assert res.hi == int(g["hi"]), f
        else:
            rj = int(g["rejection"])
            res = gpu.stack_mean_with_rejection(seq, rj, tuple(g["sig"]), int(g["normalize"]))
            assert close_enough(res.data, g["out"], frac=2e-3), f       # tiny images: allow one pixel
            if rj in (0, 1, 5):
                assert np.array_equal(res.data, g["out"]), f            # no long double involved: exact
            assert np.array_equal(res.irej, g["irej"]), f
            if int(g["normalize"]):
                assert np.allclose(res.coeff, g["coeff"], rtol=1e-11, atol=1e-9), f


# --------------------------------------------------------------------------- callbacks, cancel
def test_callbacks_and_cancel(gpu):
    fr = synth.make_frames(8, 1, 16, 16, seed=3)
    logs = []
    gpu.set_callbacks(log=logs.append)
    try:
        gpu.stack_mean_with_rejection(gpu.Sequence(fr), 2, (4, 3), 1)
        assert any("Pixel rejection in channel #0" in m for m in logs)
        assert any("Computing normalization" in m for m in logs)
        gpu.set_callbacks(should_continue=lambda: False)
        with pytest.raises(gpu.StackError) as e:
            gpu.stack_median(gpu.Sequence(fr))
        assert e.value.code == -1                                       # cancelled: -1, as the reference
    finally:
        gpu.set_callbacks()


# --------------------------------------------------------------------------- both kernels agree
@pytest.mark.parametrize("rej", list(REJ) + ["median"])
@pytest.mark.parametrize("n,norm", [(16, 0), (24, 3), (32, 2), (33, 1), (64, 4), (64, 0)])
def test_register_network_kernel_equals_general_kernel(gpu, port, rej, n, norm):
    """ssc_stack_fast.cu (register sorting network) and ssc_stack_generic.cu (shared-memory
    columns) must be interchangeable: same image, counters and guard events, both equal to the oracle."""
    fr = synth.make_frames(n, 1, 24, 64, seed=n + norm, outlier_p=8e-3, zero_p=2e-3, stars=3)
    outs = []
    for path in (0, 1, 2):                                   # automatic, general kernel, automatic with the exact normalization
        gpu.set_kernel_path(path)
        try:
            if rej == "median":
                res = gpu.stack_median(gpu.Sequence(fr), normalize=norm, want_coeff=True)
            else:
                rj, sig = REJ[rej]
                res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), rj, sig, norm)
        finally:
            gpu.set_kernel_path(0)
        outs.append(res)
    a = outs[0]
    for b in outs[1:]:
        assert np.array_equal(a.data, b.data)
        assert np.array_equal(a.irej, b.irej) and a.guard_fired == b.guard_fired and a.cap_hit == b.cap_hit
    if rej == "median":
        assert np.array_equal(a.data, port.stack_median(fr, normalize=norm, coeff=a.coeff)["out"])
    else:
        rj, sig = REJ[rej]
        same = port.stack_mean(fr, rj, sig, norm, coeff=a.coeff, sd_mode=port.SD_EXACT)
        assert np.array_equal(a.data, same["out"]) and np.array_equal(a.irej, same["irej"])
        assert a.guard_fired == same["guard_fired"]


@pytest.mark.parametrize("rej", list(REJ) + ["median"])
@pytest.mark.parametrize("n,norm", [(65, 2), (100, 3), (128, 0), (129, 4), (200, 3), (256, 1), (257, 2), (500, 3),
                                    (512, 0), (700, 4), (1024, 1)])
def test_mid_kernel_equals_general_kernel(gpu, port, rej, n, norm):
    """ssc_stack_mid.cu (65..1024 frames: register network on runs of 64 + shared-memory bitonic merges)
    against the general kernel and the oracle: same image, counters and guard events."""
    h, w = (12, 48) if n <= 256 else (6, 40)
    fr = synth.make_frames(n, 1, h, w, seed=n + norm, outlier_p=8e-3, zero_p=2e-3, stars=2)
    outs = []
    for path in (0, 1, 2):                                   # automatic (mid kernel), general kernel, exact normalization
        gpu.set_kernel_path(path)
        try:
            if rej == "median":
                res = gpu.stack_median(gpu.Sequence(fr), normalize=norm, want_coeff=True)
            else:
                rj, sig = REJ[rej]
                res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), rj, sig, norm)
        finally:
            gpu.set_kernel_path(0)
        outs.append(res)
    a = outs[0]
    for b in outs[1:]:
        assert np.array_equal(a.data, b.data)
        assert np.array_equal(a.irej, b.irej) and a.guard_fired == b.guard_fired and a.cap_hit == b.cap_hit
    if rej == "median":
        assert np.array_equal(a.data, port.stack_median(fr, normalize=norm, coeff=a.coeff)["out"])
    else:
        rj, sig = REJ[rej]
        same = port.stack_mean(fr, rj, sig, norm, coeff=a.coeff, sd_mode=port.SD_EXACT)
        assert np.array_equal(a.data, same["out"]) and np.array_equal(a.irej, same["irej"])
        assert a.guard_fired == same["guard_fired"]


@pytest.mark.parametrize("n", [8, 12, 31, 32, 33, 64, 100, 200, 257, 513, 1000])
def test_warp_kernel_linear_fit(gpu, port, n):
    """ssc_stack_warp.cu: one warp per pixel pair, every lane-count instantiation, RGB, SER order, a row
    width that leaves the last tile ragged; and columns the exact-sums screening cannot decide -- constant
    columns (sigma = 0), perfect arithmetic progressions, an exact rational tie at the high threshold --
    which must come back from the literal path with the reference's result."""
    w = 44 if n <= 257 else 20
    fr = synth.make_frames(n, 3, 5, w, seed=3 * n, outlier_p=8e-3, zero_p=2e-3, stars=1)
    fr[:, 0, 0, 0:6] = 65535                                             # saturated: sigma 0
    fr[:, 1, 1, 3] = 0
    fr[:, 2, 2, 5] = (np.arange(n) * (60000 // n)).astype(np.uint16)     # a perfect line: residuals ~ 0
    fr[:, 2, 2, 6] = 1000 + (np.arange(n) % 2)                           # two-valued
    if n == 12:                                                          # residual / sigma == 2.5 exactly in iteration 3
        fr[:, 0, 3, 7] = [1058, 1142, 1205, 1252, 1350, 1357, 1359, 1366, 1464, 1467, 1502, 1565]
    for ro in (0, 1):
        for norm in (0, 3):
            res = gpu.stack_mean_with_rejection(gpu.Sequence(fr, row_order=ro), 5, (5.0, 2.5), norm)
            same = port.stack_mean(fr, 5, (5.0, 2.5), norm, coeff=res.coeff, row_order=ro, sd_mode=port.SD_EXACT)
            assert np.array_equal(res.data, same["out"]), (ro, norm)
            assert np.array_equal(res.irej, same["irej"]) and res.guard_fired == same["guard_fired"]
    res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), 5, (-1.0, -1.0), 0)   # everything is flagged: the guard
    same = port.stack_mean(fr, 5, (-1.0, -1.0), 0, sd_mode=port.SD_EXACT)
    assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"])
    assert res.guard_fired == same["guard_fired"] > 0


@pytest.mark.parametrize("n", [513, 777, 1024])
def test_warp_kernel_median(gpu, port, n):
    fr = synth.make_random(n, 3, 4, 36, seed=n)
    for ro in (0, 1):
        res = gpu.stack_median(gpu.Sequence(fr, row_order=ro))
        assert np.array_equal(res.data, port.stack_median(fr, row_order=ro)["out"])
    fr = synth.make_frames(n, 1, 4, 36, seed=n, stars=1)
    res = gpu.stack_median(gpu.Sequence(fr), normalize=4, want_coeff=True)
    assert np.array_equal(res.data, port.stack_median(fr, normalize=4, coeff=res.coeff)["out"])


def test_mid_kernel_rgb_ser_and_fallbacks(gpu, port):
    """RGB + SER row order on the mid kernel; an odd width and registration shifts must fall back to the
    general kernel and still agree with the oracle."""
    fr = synth.make_frames(80, 3, 10, 32, seed=5, outlier_p=6e-3, zero_p=1e-3, stars=2)
    for ro in (0, 1):
        res = gpu.stack_mean_with_rejection(gpu.Sequence(fr, row_order=ro), 4, (4.0, 3.0), 3)
        same = port.stack_mean(fr, 4, (4.0, 3.0), 3, coeff=res.coeff, row_order=ro, sd_mode=port.SD_EXACT)
        assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"]), ro
    fr = synth.make_frames(70, 1, 9, 31, seed=6, outlier_p=6e-3)            # odd width
    res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), 2, (4.0, 3.0), 0)
    same = port.stack_mean(fr, 2, (4.0, 3.0), 0, sd_mode=port.SD_EXACT)
    assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"])
    fr = synth.make_frames(70, 1, 12, 32, seed=7, outlier_p=6e-3)
    sx = [(i * 7) % 5 - 2 for i in range(70)]
    sy = [(i * 3) % 5 - 2 for i in range(70)]
    res = gpu.stack_mean_with_rejection(gpu.Sequence(fr, shiftx=sx, shifty=sy), 2, (4.0, 3.0), 0)
    same = port.stack_mean(fr, 2, (4.0, 3.0), 0, shiftx=sx, shifty=sy, sd_mode=port.SD_EXACT)
    assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"])


@pytest.mark.parametrize("n", [32, 64])
def test_sigma_register_kernel_hand_backs(gpu, port, n):
    """The 32/64-frame sigma kernel keeps 3 low / 6 high candidates per pixel; pixels that reject more, and
    whole launches whose coefficients leave the fast normalization's domain, are handed to the general
    kernel through the index list.  Both must give the oracle's image and counters."""
    rng = np.random.default_rng(n)
    fr = rng.normal(3000, 30, size=(n, 1, 16, 64)).clip(0, 65535).astype(np.uint16)
    hot = rng.random(fr.shape) < 0.12                       # ~8 bright outliers per 64-frame column
    fr[hot] = rng.integers(20000, 60000, size=int(hot.sum())).astype(np.uint16)
    fr[rng.random(fr.shape) < 0.04] = 0                     # and a few dead samples (low side)
    for norm in (0, 2):
        res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), 2, (4.0, 3.0), norm)
        same = port.stack_mean(fr, 2, (4.0, 3.0), norm, coeff=res.coeff, sd_mode=port.SD_EXACT)
        assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"]), norm
        assert res.guard_fired == same["guard_fired"]
        assert res.irej[0, 1] > 6 * 16 * 64 * 0.3             # many pixels did leave the 6-candidate window
    fr2 = rng.normal(50000, 300, size=(n, 1, 16, 32)).clip(0, 65535).astype(np.uint16)
    fr2[3] = rng.integers(1, 4, size=(1, 16, 32)).astype(np.uint16)     # location ~2: mul ~ 25000
    fr2[3, 0, 2, :8] = 60000                                            # 60000 * 25000 > 2^31
    for norm in (2, 3):
        res = gpu.stack_mean_with_rejection(gpu.Sequence(fr2), 2, (4.0, 3.0), norm)
        same = port.stack_mean(fr2, 2, (4.0, 3.0), norm, coeff=res.coeff, sd_mode=port.SD_EXACT)
        assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"]), norm


@pytest.mark.parametrize("w", [36, 40, 72, 200])
def test_sigma_register_kernel_gather_variants(gpu, port, w):
    """The 32/64-frame sigma kernel gathers with 16-byte copies when rows sit on 16-byte boundaries
    (width % 8 == 0; w = 40, 72, 200: partial and several warp tiles per row) and with 4-byte copies
    otherwise (w = 36); RGB, both row orders."""
    fr = synth.make_frames(32, 3, 7, w, seed=w, outlier_p=8e-3, zero_p=2e-3, stars=2)
    for ro in (0, 1):
        res = gpu.stack_mean_with_rejection(gpu.Sequence(fr, row_order=ro), 2, (4.0, 3.0), 2)
        same = port.stack_mean(fr, 2, (4.0, 3.0), 2, coeff=res.coeff, row_order=ro, sd_mode=port.SD_EXACT)
        assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"]), (w, ro)


def test_random_shapes_all_kernels(gpu, port):
    """Seeded sweep over frame counts, image shapes (partial warp tiles, single rows, widths that are / are
    not multiples of 8), channels, row orders, rejections and normalizations: whatever kernel the
    dispatcher picks must equal the oracle bit for bit (same sigma form), counters and guard events too.
    (One of the cases has two-pixel frames with a dead pixel: IKSS scale 0, infinite coefficients, samples
    normalized to NaN -- which count as 0, as in the reference's x86-64 build.)"""
    rng = np.random.default_rng(20260101)
    rej_names = list(REJ)
    for it in range(48):
        n = int(rng.choice([2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 90, 128, 130, 257]))
        c = int(rng.choice([1, 3]))
        h = int(rng.choice([1, 2, 5, 9]))
        w = int(rng.choice([2, 6, 8, 14, 16, 24, 34, 40, 64, 66, 72, 130]))
        if n > 64 and h * w > 400:
            h = 2
        norm = int(rng.integers(0, 5))
        ro = int(rng.integers(0, 2))
        mode = rej_names[int(rng.integers(0, len(rej_names)))] if rng.random() < 0.85 else "median"
        fr = synth.make_frames(n, c, h, w, seed=1000 + it, outlier_p=1e-2, zero_p=3e-3, stars=1)
        seq = gpu.Sequence(fr, row_order=ro)
        tag = (it, n, c, h, w, norm, ro, mode)
        if mode == "median":
            res = gpu.stack_median(seq, normalize=norm, want_coeff=True)
            assert np.array_equal(res.data, port.stack_median(fr, normalize=norm, coeff=res.coeff, row_order=ro)["out"]), tag
            continue
        rj, sig = REJ[mode]
        res = gpu.stack_mean_with_rejection(seq, rj, sig, norm)
        same = port.stack_mean(fr, rj, sig, norm, coeff=res.coeff, row_order=ro, sd_mode=port.SD_EXACT)
        assert np.array_equal(res.data, same["out"]), tag
        assert np.array_equal(res.irej, same["irej"]) and res.guard_fired == same["guard_fired"], tag


def test_random_shapes_frame_walkers_and_shifts(gpu, port):
    """Seeded sweep over the whole-frame operators (sum, addmax, addmin) and the rejection mean with
    registration shifts (general kernel): shapes, channels, row orders, shifts larger than the image."""
    rng = np.random.default_rng(77)
    for it in range(40):
        n = int(rng.choice([2, 3, 7, 16, 40, 70]))
        c = int(rng.choice([1, 3]))
        h = int(rng.choice([1, 3, 8, 13]))
        w = int(rng.choice([1, 2, 7, 8, 16, 33, 64]))
        ro = int(rng.integers(0, 2))
        big = rng.random() < 0.2
        sx = rng.integers(-(w + 2) if big else -3, (w + 3) if big else 4, n).astype(np.int32)
        sy = rng.integers(-(h + 2) if big else -3, (h + 3) if big else 4, n).astype(np.int32)
        if rng.random() < 0.25:
            sx = sy = None
        fr = synth.make_random(n, c, h, w, seed=500 + it, hi=int(rng.choice([300, 5000, 65536])))
        seq = gpu.Sequence(fr, row_order=ro, shiftx=sx, shifty=sy)
        tag = (it, n, c, h, w, ro, None if sx is None else sx.tolist(), None if sy is None else sy.tolist())
        kind = int(rng.integers(0, 4))
        if kind == 0:
            res = gpu.stack_summing(seq)
            p = port.stack_sum(fr, shiftx=sx, shifty=sy, row_order=ro)
            assert np.array_equal(res.data, p["out"]) and res.hi == min(p["maxim"], 65535), tag
        elif kind in (1, 2):
            res = (gpu.stack_addmax if kind == 1 else gpu.stack_addmin)(seq)
            p = port.stack_extreme(fr, kind == 1, shiftx=sx, shifty=sy, row_order=ro)
            assert np.array_equal(res.data, p["out"]) and res.hi == p["hi"], tag
        else:
            rj, sig = REJ[list(REJ)[int(rng.integers(0, len(REJ)))]]
            res = gpu.stack_mean_with_rejection(seq, rj, sig, 0)
            p = port.stack_mean(fr, rj, sig, 0, shiftx=sx, shifty=sy, row_order=ro, sd_mode=port.SD_EXACT)
            assert np.array_equal(res.data, p["out"]) and np.array_equal(res.irej, p["irej"]), tag
            assert res.guard_fired == p["guard_fired"], tag


def test_fast_normalization_domain_fallback(gpu, port):
    """Coefficients outside the domain of the 2^52 conversion tricks (a nearly black frame against a
    bright reference: mul ~ 5e4, products beyond 2^31) take the conversion-based normalization."""
    rng = np.random.default_rng(4)
    fr = rng.normal(50000, 300, size=(8, 1, 16, 32)).clip(0, 65535).astype(np.uint16)
    fr[3] = rng.integers(1, 4, size=(1, 16, 32)).astype(np.uint16)      # location ~2: mul ~ 25000
    fr[3, 0, 2, :8] = 60000                                             # 60000 * 25000 > 2^31
    for norm in (1, 2, 3, 4):
        res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), 2, (4, 3), norm)
        assert np.isfinite(res.coeff).all()
        same = port.stack_mean(fr, 2, (4, 3), norm, coeff=res.coeff, sd_mode=port.SD_EXACT)
        assert np.array_equal(res.data, same["out"]), norm
        assert np.array_equal(res.irej, same["irej"]), norm
