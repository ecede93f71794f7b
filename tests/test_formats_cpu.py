"""Host-side logic of the file formats (no GPU): the header parsers of libsiril_stack_cuda against files laid
out by the numpy restatement in oracle/formats.py."""
import numpy as np
import pytest

import synth
from oracle import formats


def test_fits_header(pkg):
    fr = synth.make_random(1, 3, 7, 10, seed=1)[0]
    buf = formats.fits_encode(fr, exposure=12.5, extra=[("MIPS-HI", "60000"), ("MIPS-LO", "3"), ("OBSERVER", "'someone'")])
    info = pkg.fits_parse_header(buf)
    assert (info.bitpix, info.naxis, list(info.naxes)) == (16, 3, [10, 7, 3])
    assert info.bzero == 32768.0 and info.bscale == 1.0 and info.exposure == 12.5 and (info.hi, info.lo) == (60000, 3)
    assert info.data_offset == 2880 and info.data_bytes == 3 * 7 * 10 * 2 and info.frame_format == pkg.FRAMES_FITS16
    dec, cards, off = formats.fits_decode(buf)
    assert off == info.data_offset and np.array_equal(dec, fr)
    mono = formats.fits_encode(fr[:1], bitpix=8)
    info = pkg.fits_parse_header(mono)
    assert (info.bitpix, info.naxis, list(info.naxes), info.frame_format) == (8, 2, [10, 7, 1], pkg.FRAMES_RAW8)
    many = formats.fits_encode(fr[:1], extra=[("KEY%d" % i, str(i)) for i in range(40)])          # a two-block header
    assert pkg.fits_parse_header(many).data_offset == 5760
    with pytest.raises(pkg.IoError):
        pkg.fits_parse_header(buf[:800])                            # no END card
    with pytest.raises(pkg.IoError):
        pkg.fits_parse_header(formats.fits_encode(fr[:1]).replace(b"BZERO   =                32768", b"BZERO   =                    0"))


@pytest.mark.parametrize("depth,big,color", [(8, False, 0), (16, False, 0), (16, True, 0), (8, False, 100), (16, True, 101),
                                             (12, False, 8)])
def test_ser_header(pkg, depth, big, color):
    c = 3 if color in (100, 101) else 1
    fr = synth.make_random(5, c, 6, 8, seed=depth, hi=1 << min(depth, 16))
    buf = formats.ser_encode(fr, depth, big, color)
    info = pkg.ser_parse_header(buf[:178], len(buf))
    assert (info.width, info.height, info.frame_count, info.planes) == (8, 6, 5, c)
    assert info.bytes_per_sample == (1 if depth <= 8 else 2) and info.big_endian == int(big)
    assert info.frame_bytes == 8 * 6 * c * info.bytes_per_sample and info.data_offset == 178
    want = pkg.FRAMES_SER_RGB if c == 3 else (pkg.FRAMES_RAW8 if depth <= 8 else (pkg.FRAMES_RAW16_BE if big else pkg.FRAMES_U16))
    assert info.frame_format == want
    for i in range(5):
        assert np.array_equal(formats.ser_decode_frame(buf, i), fr[i])
    broken = formats.ser_encode(fr, depth, big, color, frame_count_in_header=False)             # ser_fix_broken_file
    assert pkg.ser_parse_header(broken[:178], len(broken)).frame_count == 5
    with pytest.raises(pkg.IoError):
        pkg.ser_parse_header(buf[:178], len(buf) - 10)


def test_preprocess_restatement():
    g = np.array([[100, 5, 65535, 0]], np.uint16)
    off = np.array([[10, 9, 1, 0]], np.uint16)
    flat = np.array([[2, 0, 30000, 7]], np.uint16)
    out = formats.preprocess(g, offset=off, flat=flat, level=30000.0)
    assert out.tolist() == [[65535, 0, 65534, 0]]                    # saturates both ways; flat 0 counts as 1
