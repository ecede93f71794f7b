"""Device layer (ssc_dev_*) on caller-owned device memory: band decomposition invariance (the
multi-GPU path run on one GPU, band by band), and properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev_u16(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).cuda()


def host_u16(t):
    return t.cpu().numpy().view(np.uint16)


def run_pipeline(gpu, frames_dev, n, c, h, w, bands, rejection, sig, normalize, median=False):
    """The per-rank sequence of bench.py, band by band on one GPU; histograms add up across bands."""
    sp = torch.cuda.current_stream().cuda_stream
    out = torch.zeros((c, h, w), dtype=torch.int16, device="cuda")
    hist = torch.zeros((n, gpu.HIST_STRIDE), dtype=torch.int32, device="cuda")
    scratch = torch.empty((n, 65536), dtype=torch.int32, device="cuda")
    stats = torch.empty((4, n), dtype=torch.float64, device="cuda")
    coeff = torch.empty((3, n), dtype=torch.float64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    counters = torch.zeros(gpu.N_COUNTERS, dtype=torch.int64, device="cuda")
    mk = lambda r0, rn: gpu.dev.band(frames_dev.data_ptr(), n, c, h, w, out_row0=r0, out_rows=rn)
    if normalize:
        for r0, rn in bands:
            gpu.dev.histogram(mk(r0, rn), hist.data_ptr(), sp)
        gpu.dev.ikss(hist.data_ptr(), n, stats.data_ptr(), scratch.data_ptr(), sp)
        gpu.dev.norm_coeffs(stats.data_ptr(), n, 0, normalize, coeff.data_ptr(), status.data_ptr(), sp)
    for r0, rn in bands:
        if median:
            gpu.dev.stack_median(mk(r0, rn), out.data_ptr(), normalize, coeff.data_ptr(), out_layer_stride=h * w,
                                 out_buf_row0=0, stream=sp)
        else:
            gpu.dev.stack_mean(mk(r0, rn), out.data_ptr(), counters.data_ptr(), rejection, sig, normalize,
                               coeff.data_ptr(), out_layer_stride=h * w, out_buf_row0=0, stream=sp)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    return host_u16(out), counters.cpu().numpy(), coeff.cpu().numpy()


@pytest.mark.parametrize("n", [24, 32, 70])       # block-queue kernel, register sigma kernel (full stack), mid kernel
@pytest.mark.parametrize("rej,sig", [(2, (4.0, 3.0)), (4, (4.0, 3.0)), (5, (5.0, 2.5))])
def test_band_decomposition_is_invariant(gpu, port, rej, sig, n):
    c, h, w = 3, 45, 64
    fr = synth.make_frames(n, c, h, w, seed=31, outlier_p=5e-3, zero_p=1e-3, stars=4)
    d = dev_u16(fr)
    whole = run_pipeline(gpu, d, n, c, h, w, [(0, h)], rej, sig, 3)
    for bands in ([(0, 23), (23, 22)], [(0, 6), (6, 6), (12, 6), (18, 6), (24, 6), (30, 5), (35, 5), (40, 5)]):
        part = run_pipeline(gpu, d, n, c, h, w, bands, rej, sig, 3)
        assert np.array_equal(part[0], whole[0])
        assert np.array_equal(part[1], whole[1])
        assert np.array_equal(part[2], whole[2])                        # statistics exact under sharding
    ref = port.stack_mean(fr, rej, sig, 3, coeff=whole[2], sd_mode=port.SD_EXACT)
    assert np.array_equal(whole[0], ref["out"])
    assert np.array_equal(whole[1][:6].reshape(3, 2), ref["irej"].astype(np.int64))


def test_synthetic_generator_is_band_independent(gpu):
    n, c, h, w = 3, 3, 64, 96
    sp = torch.cuda.current_stream().cuda_stream
    full = torch.empty((n, c, h, w), dtype=torch.int16, device="cuda")
    gpu.dev.synth_frames(full.data_ptr(), n, c, h, w, 0, 0, h, 0, 1234, sp)
    part = torch.empty((n, c, 20, w), dtype=torch.int16, device="cuda")
    gpu.dev.synth_frames(part.data_ptr(), n, c, 20, w, 0, 17, h, 0, 1234, sp)
    torch.cuda.synchronize()
    assert torch.equal(full[:, :, 17:37, :], part)
    later = torch.empty((1, c, h, w), dtype=torch.int16, device="cuda")
    gpu.dev.synth_frames(later.data_ptr(), 1, c, h, w, 2, 0, h, 0, 1234, sp)
    torch.cuda.synchronize()
    assert torch.equal(full[2:3], later)


# ---------------------------------------------------------------------------------------------
# BASELINE.json's full sizes, checked through size-independent properties
# ---------------------------------------------------------------------------------------------
def test_cfg2_median_full_size(gpu):
    """configs[1]: median of 32 x 4096x4096 mono.  Independent check: torch.sort along the frame
    axis on the same device buffer, floor((s[15] + s[16]) / 2) (D7)."""
    n, c, h, w = 32, 1, 4096, 4096
    sp = torch.cuda.current_stream().cuda_stream
    fr = torch.empty((n, c, h, w), dtype=torch.int16, device="cuda")
    gpu.dev.synth_frames(fr.data_ptr(), n, c, h, w, 0, 0, h, 0, 0x5349524C, sp)
    out = torch.empty((c, h, w), dtype=torch.int16, device="cuda")
    gpu.dev.stack_median(gpu.dev.band(fr.data_ptr(), n, c, h, w), out.data_ptr(), stream=sp)
    torch.cuda.synchronize()
    got = out.view(-1).to(torch.int32) & 0xFFFF
    x = (fr.view(n, -1).to(torch.int32) & 0xFFFF)
    step = 4096 * 512                                                   # sort in slabs to bound memory
    for i in range(0, x.shape[1], step):
        s = torch.sort(x[:, i:i + step], dim=0).values
        assert torch.equal((s[15] + s[16]) // 2, got[i:i + step])


def test_cfg1_sum_full_size(gpu):
    """configs[0]: sum of 16 x 1024x1024 mono; independent check in torch int64 (+ pixel 0 quirk, rescale)."""
    n, c, h, w = 16, 1, 1024, 1024
    sp = torch.cuda.current_stream().cuda_stream
    fr = torch.empty((n, c, h, w), dtype=torch.int16, device="cuda")
    gpu.dev.synth_frames(fr.data_ptr(), n, c, h, w, 0, 0, h, 0, 0x5349524C, sp)
    acc = torch.empty((c, h, w), dtype=torch.int32, device="cuda")
    mx = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.empty((c, h, w), dtype=torch.int16, device="cuda")
    gpu.dev.sum_accumulate(gpu.dev.band(fr.data_ptr(), n, c, h, w), acc.data_ptr(), mx.data_ptr(), sp)
    gpu.dev.sum_finish(acc.data_ptr(), c * h * w, mx.data_ptr(), out.data_ptr(), sp)
    torch.cuda.synchronize()
    s = (fr.to(torch.int64) & 0xFFFF).sum(0).view(-1)
    s[0] = 0
    m = int(s.max())
    assert int(mx.item()) == m and m > 65535
    expect = torch.floor(s.to(torch.float64) * (65535.0 / m) + 0.5).clamp(0, 65535).to(torch.int64)
    assert torch.equal(expect, out.view(-1).to(torch.int64) & 0xFFFF)


def test_cfg3_sigma_full_width_rows_against_oracle(gpu, port):
    """configs[2] shape per row (64 RGB frames, 6000 px wide, sigma clip, multiplicative
    normalization) on a 96-row slab; 12 full rows are replayed through the oracle."""
    n, c, h, w = 64, 3, 96, 6000
    sp = torch.cuda.current_stream().cuda_stream
    fr = torch.empty((n, c, h, w), dtype=torch.int16, device="cuda")
    gpu.dev.synth_frames(fr.data_ptr(), n, c, h, w, 0, 1000, 4000, 0, 0x5349524C, sp)
    out, counters, coeff = run_pipeline(gpu, fr, n, c, h, w, [(0, h)], 2, (4.0, 3.0), 2)
    rows = list(range(0, h, 8))
    sub = host_u16(fr[:, :, rows, :].contiguous())
    same = port.stack_mean(sub, 2, (4.0, 3.0), 2, coeff=coeff, sd_mode=port.SD_EXACT)
    assert np.array_equal(out[:, rows, :], same["out"])
    gsl = port.stack_mean(sub, 2, (4.0, 3.0), 2, coeff=coeff, sd_mode=port.SD_GSL)
    d = np.abs(out[:, rows, :].astype(np.int32) - gsl["out"].astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-4                    # the bar of BASELINE.json
    assert np.array_equal(same["irej"], gsl["irej"])
    assert counters[6] == 0 and counters[7] == 0                       # no guard event, no iteration cap
    # idempotence-type property at this size: stacking 64 copies of one frame returns that frame
    one = fr[:1].expand(n, c, h, w).contiguous()
    out1, cnt1, _ = run_pipeline(gpu, one, n, c, h, w, [(0, h)], 2, (4.0, 3.0), 0)
    assert np.array_equal(out1, host_u16(fr[0])) and cnt1[:6].sum() == 0


def test_non_finite_coefficients_are_deterministic(gpu):
    """A frame whose IKSS scale is 0 gets infinite / NaN coefficients (stacking.c:95-121).  Samples
    normalized to NaN become 0 (what the reference's x86-64 build does), +inf becomes 65535 -- on every
    kernel path."""
    import torch
    dev = torch.device("cuda", 0)
    sp = torch.cuda.current_stream().cuda_stream
    raw = np.array([[[[100, 0, 5000, 65535, 7, 8, 9, 10]]], [[[200, 300, 400, 500, 600, 700, 800, 900]]]], dtype=np.uint16)
    fr = torch.tensor(raw.view(np.int16), device=dev)
    n, c, h, w = raw.shape
    cases = [([np.inf, 0.0], [1.0, 1.0], [np.inf, 1.0]), ([np.nan, 0.0], [1.0, 1.0], [np.nan, 1.0]),
             ([0.0, 0.0], [np.inf, 1.0], [1.0, 1.0])]
    for off, mul, sc in cases:
        for norm in (1, 2, 3, 4):
            if norm in (1, 2):
                sc = [1.0, 1.0]         # compute_normalization leaves scale at 1.0 in these modes (stacking.c:130-134)
            with np.errstate(all="ignore"):
                x = raw[:, 0, 0, :].astype(np.float64)
                if norm in (1, 3):      # stacking.c:1640-1644
                    y = x * np.array(sc)[:, None] - np.array(off)[:, None]
                else:                   # stacking.c:1646-1650
                    y = (x * np.array(sc)[:, None]) * np.array(mul)[:, None]
                v = np.where(np.isnan(y) | (y <= 0), 0.0, np.where(y > 65535.0, 65535.0, np.floor(y + 0.5)))
                expect = np.floor(v.sum(0) / n + 0.5).clip(0, 65535).astype(np.uint16)
            coeff = torch.tensor(np.array([off, mul, sc], dtype=np.float64), device=dev)
            for path in (0, 1):
                gpu.set_kernel_path(path)
                try:
                    out = torch.zeros((c, h, w), dtype=torch.int16, device=dev)
                    cnt = torch.zeros(8, dtype=torch.int64, device=dev)
                    gpu.dev.stack_mean(gpu.dev.band(fr.data_ptr(), n, c, h, w), out.data_ptr(), cnt.data_ptr(), 0,
                                       (0.0, 0.0), norm, coeff.data_ptr(), stream=sp)
                    torch.cuda.synchronize()
                finally:
                    gpu.set_kernel_path(0)
                assert np.array_equal(out.cpu().numpy().view(np.uint16).ravel(), expect), (off, mul, sc, norm, path)


@pytest.mark.parametrize("n,w", [(200, 32), (100, 48), (300, 16), (64, 64)])
@pytest.mark.parametrize("norm", [1, 2, 3, 4])
def test_one_fma_normalization_is_only_used_where_it_is_exact(gpu, port, n, w, norm):
    """The 65..1024-frame kernel (ssc_stack_mid.cu) normalizes with one FMA per sample on launches whose coefficients
    k_verify_cheap_norm cleared (all 65536 sample values of every frame equal to the reference's individually
    rounded form, stacking.c:1635-1652 + round_to_WORD) and with the reference's form otherwise.  Coefficient sets
    that hit exact .5 ties (where nearest-even and round-half-up part), negative, tiny and huge factors must
    therefore give the oracle's image on every path -- and so must ordinary ones."""
    import torch
    dev = torch.device("cuda", 0)
    sp = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(1000 * n + norm)
    c, h = 1, 8
    raw = rng.integers(0, 65536, size=(n, c, h, w), dtype=np.uint16)
    raw[:, :, 0, :8] = np.array([0, 1, 2, 3, 65535, 65534, 32768, 32767], dtype=np.uint16)
    fr = torch.tensor(raw.view(np.int16), device=dev)
    sets = {
        "ordinary": (rng.uniform(-300, 300, n), rng.uniform(0.8, 1.25, n), rng.uniform(0.8, 1.25, n)),
        "ties": (rng.choice([0.5, -0.5, 1.5, 100.5, -0.25], n), rng.choice([0.5, 1.5, 0.75, 1.25, 1.0], n),
                 rng.choice([0.5, 1.5, 2.5, 1.0, 0.25], n)),
        "one bad frame": (np.r_[rng.uniform(-300, 300, n - 1), 0.5], np.r_[rng.uniform(0.8, 1.25, n - 1), 0.5],
                          np.r_[rng.uniform(0.8, 1.25, n - 1), 1.5]),
        "negative": (rng.uniform(-300, 300, n), -rng.uniform(0.8, 1.25, n), rng.uniform(0.8, 1.25, n)),
        "tiny and huge": (rng.uniform(-3, 3, n), rng.choice([1e-9, 3e-5, 7.0, 900.0, 2e4], n), rng.choice([1e-7, 0.013, 33.0], n)),
    }
    for name, (off, mul, sc) in sets.items():
        if norm in (1, 2):
            sc = np.ones(n)            # compute_normalization leaves scale at 1.0 in these modes (stacking.c:130-134)
        coeff_h = np.array([off, mul, sc], dtype=np.float64)
        coeff = torch.tensor(coeff_h, device=dev)
        want = port.stack_mean(raw, 2, (4.0, 3.0), norm, coeff=coeff_h, sd_mode=port.SD_EXACT)
        for path in (0, 2, 1):
            gpu.set_kernel_path(path)
            try:
                out = torch.zeros((c, h, w), dtype=torch.int16, device=dev)
                cnt = torch.zeros(8, dtype=torch.int64, device=dev)
                gpu.dev.stack_mean(gpu.dev.band(fr.data_ptr(), n, c, h, w), out.data_ptr(), cnt.data_ptr(), 2,
                                   (4.0, 3.0), norm, coeff.data_ptr(), stream=sp)
                torch.cuda.synchronize()
            finally:
                gpu.set_kernel_path(0)
            assert np.array_equal(out.cpu().numpy().view(np.uint16), want["out"]), (name, norm, path)
            assert np.array_equal(cnt.cpu().numpy()[:6].reshape(3, 2), want["irej"]), (name, norm, path)
