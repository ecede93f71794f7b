"""Hand-made known-answer tests of the restated semantics:
the reference has no tests of its own, so these fix the quirks in writing."""
import numpy as np
import pytest

import synth


def col(port, values, rej, sig, sd_mode=0):
    return port.reject_column(values, rej, sig, sd_mode)


def test_round_to_word(port):
    f = port.lib().so_round_to_word
    assert [f(-1.0), f(0.0), f(0.49), f(0.5), f(1.5), f(65534.5), f(65535.0), f(65535.4), f(65536.0), f(1e9)] == \
        [0, 0, 0, 1, 2, 65535, 65535, 65535, 65535, 65535]


def test_median_even_truncates(port):
    fr = np.array([3, 4, 10, 20], np.uint16).reshape(4, 1, 1, 1)
    assert port.stack_median(fr)["out"][0, 0, 0] == 7          # (4+10)/2 = 7.0
    fr = np.array([3, 4, 9, 20], np.uint16).reshape(4, 1, 1, 1)
    assert port.stack_median(fr)["out"][0, 0, 0] == 6          # 6.5 truncated, not rounded (D7)
    fr = np.array([9, 3, 20], np.uint16).reshape(3, 1, 1, 1)
    assert port.stack_median(fr)["out"][0, 0, 0] == 9


def test_sum_pixel0_and_rescale(port):
    fr = np.full((4, 1, 2, 4), 100, np.uint16)
    r = port.stack_sum(fr)
    assert r["out"].ravel().tolist() == [0, 400, 400, 400, 400, 400, 400, 400]     # D8
    fr = np.full((4, 1, 2, 4), 30000, np.uint16)
    fr[:, 0, 1, 1] = 60000
    r = port.stack_sum(fr)
    assert r["maxim"] == 240000
    ratio = 65535.0 / 240000
    assert r["out"][0, 1, 1] == 65535 and r["out"][0, 0, 1] == int(120000 * ratio + 0.5)
    assert r["out"][0, 0, 0] == 0


def test_no_rejection_mean_rounds_half_up(port):
    assert col(port, [1, 2], 0, (0, 0))[0] == 2               # 1.5 -> 2
    assert col(port, [1, 2, 2], 0, (0, 0))[0] == 2            # 1.667
    assert col(port, [1, 1, 2, 1], 0, (0, 0))[0] == 1         # 1.25


def test_sigma_clip_planted_outlier(port):
    base = [1000 + (i * 7) % 13 for i in range(20)]
    v, rej, guard, cap = col(port, base + [30000], 2, (4.0, 3.0))
    assert rej.tolist() == [0, 1] and guard == 0
    assert v == int(sum(base) / 20.0 + 0.5)
    v, rej, guard, cap = col(port, base + [0], 2, (4.0, 3.0))
    assert rej.tolist() == [1, 0]
    assert v == int(sum(base) / 20.0 + 0.5)


def test_percentile_clip(port):
    vals = [100] * 9 + [50, 200]
    # median 100: (100-50)/100 = .5 > .2 low ; (200-100)/100 = 1 > .1 high
    v, rej, _, _ = col(port, vals, 1, (0.2, 0.1))
    assert rej.tolist() == [1, 1] and v == 100
    # median 0: every positive sample is "high" (x/0 = inf); all flagged -> the last one survives
    v, rej, _, _ = col(port, [0, 0, 0, 5], 1, (0.2, 0.1))
    assert rej.tolist() == [0, 1] and v == 0


def test_sigmedian_replaces_instead_of_removing(port):
    base = [1000 + (i * 7) % 13 for i in range(20)]
    v, rej, _, _ = col(port, base + [30000], 3, (4.0, 3.0))
    s = sorted(base + [30000])
    med = int((s[10]) + 0.5)
    assert rej[1] >= 1
    assert v == int((sum(base) + med) / 21.0 + 0.5)           # N stays 21: the outlier became the median


def test_winsorized_and_linearfit_drop_outlier(port):
    base = [1000 + (i * 7) % 13 for i in range(30)]
    for rej_mode, sig in ((4, (4.0, 3.0)), (5, (5.0, 2.5))):
        v, rej, guard, _ = col(port, base + [40000], rej_mode, sig)
        assert rej[1] >= 1 and guard == 0
        assert abs(v - sum(base) / 30.0) <= 2


def test_guard_keeps_at_least_some_samples(port):
    # 6 samples: anything that would leave <= 4 stops the flagging (stacking.c:1684)
    v, rej, guard, _ = col(port, [10, 10, 10, 10, 5000, 9000], 2, (0.5, 0.5))
    assert guard in (0, 1)
    assert 10 <= v <= 9000


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_coefficient_algebra(port, mode):
    loc = np.array([1000.0, 1100.0, 950.0])
    sc = np.array([40.0, 44.0, 38.0])
    c = port.norm_coeffs(mode, loc, sc, ref_image=0)
    off, mul, scale = c
    if mode == 1:
        assert np.array_equal(off, 1.0 * loc - loc[0]) and (mul == 1).all() and (scale == 1).all()
    if mode == 3:
        assert np.array_equal(scale, sc[0] / sc) and np.array_equal(off, scale * loc - loc[0])
    if mode == 2:
        assert np.array_equal(mul, loc[0] / loc) and (off == 0).all() and (scale == 1).all()
    if mode == 4:
        assert np.array_equal(mul, loc[0] / loc) and np.array_equal(scale, sc[0] / sc)


def test_ikss_on_constant_and_empty(port):
    hist = np.zeros(65536, np.uint32)
    assert port.ikss_hist(hist)[:2] == (0.0, 0.0)              # no sample: location = scale = 0
    hist[1234] = 500
    loc, sc, it = port.ikss_hist(hist)
    assert abs(loc - 1234.0) < 1e-9 and sc == 0.0              # mad = 0 -> bwmv = 0 -> scale 0


def test_fit_formulations_agree(port):
    """LINEARFIT: the straight line from exact integer sums (what the warp kernel screens with,
    SO_FIT_EXACT_SUMS) against gsl_fit_linear's literal running means.  Away from exact ties of a
    residual with a threshold both flag the same samples: identical images and counters on seeded
    stacks; the one known tie (residual / sigma == 2.5 exactly in the third iteration) is the kind of
    pixel the kernel hands back to the literal path."""
    for n, seed in ((24, 1), (64, 2), (200, 3), (1000, 4)):
        fr = synth.make_frames(n, 1, 12, 96, seed=seed, outlier_p=5e-3, zero_p=1e-3, stars=3, planetary=(n == 1000))
        lit = port.stack_mean(fr, port.LINEARFIT, (5.0, 2.5), sd_mode=port.SD_EXACT)
        exa = port.stack_mean(fr, port.LINEARFIT, (5.0, 2.5), sd_mode=port.SD_EXACT | port.FIT_EXACT)
        assert np.array_equal(lit["out"], exa["out"]), n
        assert np.array_equal(lit["irej"], exa["irej"]) and lit["guard_fired"] == exa["guard_fired"]
    tie = [1058, 1142, 1205, 1252, 1350, 1357, 1359, 1366, 1464, 1467, 1502, 1565]
    a = port.reject_column(tie, port.LINEARFIT, (5.0, 2.5), port.SD_EXACT)
    b = port.reject_column(tie, port.LINEARFIT, (5.0, 2.5), port.SD_EXACT | port.FIT_EXACT)
    assert a[0] != b[0]                                      # rounding noise decides an exact tie


# --------------------------------------------------------------------------- the one-FMA normalization
def _cheap_form(p, mode, off, mul, sc):
    """csrc/ssc_network.cuh:cheap_sample restated with exact rationals (an FMA rounds once: float(Fraction) is the
    correctly rounded quotient of two integers).  Returns the WORD the kernel would produce."""
    from fractions import Fraction as F
    two52, magic = 4503599627370496.0, 6755399441055744.0
    d = F(two52) + p                                                    # the register pair {p, 0x43300000}
    if mode in (2, 4):                                                  # multiplicative kinds
        a = float(np.float64(sc) * np.float64(mul)) if mode == 4 else mul
        c = float(F(-two52) * F(a) + F(two52))                          # fma(-2^52, a, 2^52)
        u = float(d * F(a) + F(c))                                      # fma(2^52 + p, a, c) = 2^52 + RN_int(p a)
        v = int(u - two52)
    else:                                                               # additive kinds
        a = sc if mode == 3 else 1.0
        c = float(np.float64(-two52) * np.float64(a))
        t = float(d * F(a) + F(c))                                      # = RN(p a)
        u = float(np.float64(t) - np.float64(off)) + magic              # (t - off) + 1.5 2^52, each rounded
        v = int(u - magic)
    return min(max(v, 0), 65535)


def _reference_form(p, mode, off, mul, sc):
    x = np.float64(p)
    y = x * np.float64(sc) - np.float64(off) if mode in (1, 3) else (x * np.float64(sc)) * np.float64(mul)   # stacking.c:1640-1650
    if not y > 0:
        return 0
    return 65535 if y > 65535.0 else int(y + np.float64(0.5))           # round_to_WORD, utils.c:68-74


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_one_fma_normalization_model(mode):
    """The kernels' one-FMA form of the normalization equals the reference's individually rounded form for ordinary
    coefficients and differs from it at exact .5 ties (nearest-even against round-half-up) -- which is why the library
    only uses it on launches where k_verify_cheap_norm found all 65536 sample values of every frame equal
    (tests/test_gpu_device_layer.py runs that on the device)."""
    rng = np.random.default_rng(mode)
    samples = np.r_[0:300, 32760:32776, 65000:65536, rng.integers(0, 65536, 1500)]
    for _ in range(3):                                                  # ratios of IKSS estimates: no ties in reach
        off, mul, sc = rng.uniform(-300, 300), rng.uniform(0.8, 1.25), rng.uniform(0.8, 1.25)
        if mode in (1, 2):
            sc = 1.0                                                    # stacking.c:130-134
        assert all(_cheap_form(int(p), mode, off, mul, sc) == _reference_form(int(p), mode, off, mul, sc) for p in samples)
    off, mul, sc = 0.5, 0.5, (1.0 if mode in (1, 2) else 0.5)           # half-integer products on every other sample
    differ = [int(p) for p in range(1, 64) if _cheap_form(p, mode, off, mul, sc) != _reference_form(p, mode, off, mul, sc)]
    assert differ, "ties must separate the two forms (otherwise the verification would be pointless)"
