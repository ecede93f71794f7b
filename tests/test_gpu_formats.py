"""The frames as they lie in the files: the operators read FITS data units and SER frames
straight from the file's bytes, decode them on the GPU behind the copy, run preprocess on them and stack; the
result must be what stacking the decoded (numpy restatement, oracle/formats.py) frames gives.  And savefits."""
import ctypes as C
import os

import numpy as np
import pytest

import synth
from oracle import formats

pytestmark = pytest.mark.gpu


def addr(buf, offset):
    return C.addressof(C.c_char.from_buffer(buf)) + offset


def fits_sequence(gpu, frames, bitpix=16, **kw):
    """One FITS file image per frame, in memory; the sequence points at their data units."""
    files = [bytearray(formats.fits_encode(f, bitpix=bitpix)) for f in frames]
    infos = [gpu.fits_parse_header(bytes(b)) for b in files]
    n, (c, h, w) = len(frames), frames[0].shape
    assert all((i.naxes[0], i.naxes[1], i.naxes[2]) == (w, h, c) for i in infos)
    seq = gpu.Sequence([addr(b, i.data_offset) for b, i in zip(files, infos)], shape=(n, c, h, w),
                       frame_format=infos[0].frame_format, **kw)
    seq._files = files
    return seq


@pytest.mark.parametrize("c", [1, 3])
def test_fits16_frames_stack_like_decoded_frames(gpu, port, c, monkeypatch):
    n, h, w = 12, 33, 40
    fr = synth.make_frames(n, c, h, w, seed=5 + c, outlier_p=5e-3, zero_p=1e-3, stars=3)
    fr[0, 0, 0, :4] = [0, 1, 32768, 65535]
    native = gpu.Sequence(fr)
    for chunk in (None, 4 * n * w * 2):                     # one chunk, and many small ones (odd sample counts too)
        if chunk:
            monkeypatch.setenv("SSC_CHUNK_BYTES", str(chunk))
        seq = fits_sequence(gpu, fr)
        for run in (lambda s: gpu.stack_mean_with_rejection(s, 2, (4.0, 3.0), 3), lambda s: gpu.stack_median(s),
                    lambda s: gpu.stack_summing(s), lambda s: gpu.stack_addmax(s)):
            a, b = run(native), run(seq)
            assert np.array_equal(a.data, b.data) and np.array_equal(a.irej, b.irej) and a.hi == b.hi
            if a.coeff is not None:
                assert np.array_equal(a.coeff, b.coeff)
    monkeypatch.delenv("SSC_CHUNK_BYTES")
    shifted = fits_sequence(gpu, fr, shiftx=[i % 5 - 2 for i in range(n)], shifty=[(3 * i) % 7 - 3 for i in range(n)])
    ref = gpu.Sequence(fr, shiftx=[i % 5 - 2 for i in range(n)], shifty=[(3 * i) % 7 - 3 for i in range(n)])
    a, b = gpu.stack_mean_with_rejection(ref, 4, (4.0, 3.0), 1), gpu.stack_mean_with_rejection(shifted, 4, (4.0, 3.0), 1)
    assert np.array_equal(a.data, b.data) and np.array_equal(a.irej, b.irej)


def test_fits8_frames(gpu, port):
    fr = synth.make_random(6, 3, 17, 24, seed=2, hi=256)
    a = gpu.stack_mean_with_rejection(gpu.Sequence(fr), 0, (0, 0), 3)
    b = gpu.stack_mean_with_rejection(fits_sequence(gpu, fr, bitpix=8), 0, (0, 0), 3)
    assert np.array_equal(a.data, b.data) and np.array_equal(a.coeff, b.coeff)          # hist_size 256 on both routes


@pytest.mark.parametrize("depth,big,color", [(8, False, 0), (16, False, 0), (16, True, 0), (8, False, 100), (16, False, 100),
                                             (16, True, 101), (8, False, 101)])
def test_ser_frames_stack_like_decoded_frames(gpu, port, depth, big, color, monkeypatch):
    """A whole SER file in memory: header parsed by the library, frames decoded on the device (8 -> 16 bit, byte
    order, RGB / BGR planes), SER row order."""
    c = 3 if color in (100, 101) else 1
    n, h, w = 9, 21, 32
    fr = synth.make_frames(n, c, h, w, seed=depth + color, outlier_p=5e-3, stars=2)
    if depth == 8:
        fr = (fr >> 8).astype(np.uint16)
    buf = bytearray(formats.ser_encode(fr, depth, big, color))
    info = gpu.ser_parse_header(bytes(buf[:178]), len(buf))
    decoded = np.stack([formats.ser_decode_frame(bytes(buf), i) for i in range(n)])
    assert np.array_equal(decoded, fr)
    native = gpu.Sequence(decoded, row_order=gpu.ROWS_SER)
    for chunk in (None, 3 * n * w * 2):
        if chunk:
            monkeypatch.setenv("SSC_CHUNK_BYTES", str(chunk))
        seq = gpu.Sequence([addr(buf, info.data_offset + i * info.frame_bytes) for i in range(info.frame_count)],
                           shape=(n, c, h, w), row_order=gpu.ROWS_SER, frame_format=info.frame_format,
                           ser_bytes_per_sample=info.bytes_per_sample, ser_big_endian=info.big_endian, ser_bgr=int(color == 101))
        for run in (lambda s: gpu.stack_mean_with_rejection(s, 2, (4.0, 3.0), 2), lambda s: gpu.stack_median(s),
                    lambda s: gpu.stack_summing(s)):
            a, b = run(native), run(seq)
            assert np.array_equal(a.data, b.data) and np.array_equal(a.irej, b.irej) and a.hi == b.hi
    monkeypatch.delenv("SSC_CHUNK_BYTES")



@pytest.mark.parametrize("use", [("offset",), ("dark", "flat"), ("offset", "dark", "flat")])
def test_preprocess_fused_into_the_load(gpu, port, use, monkeypatch):
    """preprocess (siril.c:945-962) on the decoded rows before the statistics and the stack see them: equals
    stacking frames calibrated by the numpy restatement -- native and FITS frames, RGB, chunked, two pieces."""
    n, c, h, w = 10, 3, 29, 40
    fr = synth.make_frames(n, c, h, w, seed=31, outlier_p=5e-3, zero_p=1e-3, stars=3)
    rng = np.random.default_rng(7)
    masters = dict(offset=rng.integers(90, 110, (c, h, w)).astype(np.uint16), dark=rng.integers(0, 400, (c, h, w)).astype(np.uint16),
                   flat=rng.integers(20000, 40000, (c, h, w)).astype(np.uint16))
    masters["flat"][0, 3, 5] = 0                                         # fdiv: 0 counts as 1
    kw = {("pp_" + k): masters[k] for k in use}
    level = 30000.0
    cal = np.stack([formats.preprocess(f, masters["offset"] if "offset" in use else None, masters["dark"] if "dark" in use else None,
                                       masters["flat"] if "flat" in use else None, level) for f in fr])
    want = gpu.stack_mean_with_rejection(gpu.Sequence(cal), 2, (4.0, 3.0), 3)
    same = port.stack_mean(cal, 2, (4.0, 3.0), 3, coeff=want.coeff, sd_mode=port.SD_EXACT)
    assert np.array_equal(want.data, same["out"])
    for mk in (lambda: gpu.Sequence(fr, pp_flat_level=level, **kw), lambda: fits_sequence(gpu, fr, pp_flat_level=level, **kw)):
        for env in ({}, {"SSC_CHUNK_BYTES": str(5 * n * w * 2)}, {"SSC_HBM_BUDGET_BYTES": str(16 * (n * c * w * 2 + c * w * 10))}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = gpu.stack_mean_with_rejection(mk(), 2, (4.0, 3.0), 3)
            for k in env:
                monkeypatch.delenv(k)
            assert np.array_equal(got.coeff, want.coeff), env
            assert np.array_equal(got.data, want.data) and np.array_equal(got.irej, want.irej), env


def test_savefits_round_trip(gpu, tmp_path):
    """ssc_savefits: a file that reads back (numpy restatement of cfitsio's TUSHORT read, and the library's own
    parser + device decoder) as the image that was saved; header cards and 2880-byte blocks as the standard asks."""
    import torch
    fr = synth.make_frames(8, 3, 19, 26, seed=3, stars=2)
    res = gpu.stack_mean_with_rejection(gpu.Sequence(fr, exposure=[1.5] * 8), 2, (4.0, 3.0), 1)
    path = os.path.join(tmp_path, "stack.fit")
    gpu.savefits(path, res.data, hi=int(res.data.max()), lo=int(res.data.min()), exposure=res.exposure)
    buf = open(path, "rb").read()
    assert len(buf) % 2880 == 0
    dec, cards, off = formats.fits_decode(buf)
    assert np.array_equal(dec, res.data)
    assert cards["BITPIX"] == "16" and cards["BZERO"] == "32768" and cards["BSCALE"] == "1" and cards["NAXIS"] == "3"
    assert float(cards["EXPTIME"]) == 12.0 and int(cards["MIPS-HI"]) == int(res.data.max())
    info = gpu.fits_parse_header(buf)
    raw = torch.frombuffer(bytearray(buf[info.data_offset:info.data_offset + info.data_bytes]), dtype=torch.int16).cuda()
    gpu._lib.ssc_dev_decode_fits16(raw.data_ptr(), raw.numel(), None)
    torch.cuda.synchronize()
    assert np.array_equal(raw.cpu().numpy().view(np.uint16).reshape(res.data.shape), res.data)
    with pytest.raises(gpu.IoError):
        gpu.savefits(path, res.data, overwrite=False)
    mono = res.data[:1]
    gpu.savefits(path, mono)
    dec, cards, _ = formats.fits_decode(open(path, "rb").read())
    assert np.array_equal(dec, mono) and cards["NAXIS"] == "2" and "NAXIS3" not in cards


def test_post_stack_statistics(gpu):
    """ssc_image_statistics = statistics(STATS_BASIC, nullcheck) per layer of the stacked image (what noise() and the
    stacking summary compute): counts, min, max and the histogram median exact; mean, sigma and the noise estimate
    against the float64 restatement."""
    fr = synth.make_frames(10, 3, 64, 96, seed=17, outlier_p=2e-3, zero_p=2e-3, stars=6)
    res = gpu.stack_mean_with_rejection(gpu.Sequence(fr), 2, (4.0, 3.0), 1)
    img = res.data.copy()
    img[1, 5, :] = 0                                          # a dead row: no differences at all
    img[2, 7, 3:] = 0                                         # a row with a single difference
    st = gpu.image_statistics(img)
    for c in range(3):
        want = formats.basic_statistics(img[c])
        for k in ("total", "ngoodpix", "median", "min", "max"):
            assert st[c][k] == want[k], (c, k)
        for k in ("mean", "sigma", "bgnoise"):
            assert abs(st[c][k] - want[k]) <= 1e-12 * abs(want[k]), (c, k, st[c][k], want[k])
        assert st[c]["normValue"] == 65535.0
    low = (img[:1] >> 8).astype(np.uint16)
    assert gpu.image_statistics(low)[0]["normValue"] == 255.0
    with pytest.raises(gpu.IoError):
        gpu.image_statistics(np.zeros((1, 8, 8), np.uint16))           # "Error: no data computed."
    st0 = gpu.image_statistics(low, nullcheck=False)[0]
    w0 = formats.basic_statistics(low[0], nullcheck=False)
    assert st0["ngoodpix"] == low.size and st0["median"] == w0["median"] and st0["min"] == w0["min"]
