"""numpy restatement of how Siril 0.9.7 reads and writes pixel data (TEST INFRASTRUCTURE ONLY): FITS
BITPIX 16 / BZERO 32768 and BITPIX 8 (src/io/image_format_fits.c through cfitsio's TUSHORT / TBYTE
conversion), SER frames (src/io/ser.c) and preprocess (src/core/siril.c).  cfitsio is not part of
/root/reference: the conversion it applies (physical = BSCALE * stored + BZERO, stored big-endian) is the
FITS standard's, restated here."""
import struct

import numpy as np


def _card(key, value, comment=""):
    return ("%-8s= %20s / %-47s" % (key, value, comment))[:80].ljust(80)


def fits_encode(frame, exposure=0.0, bitpix=16, extra=()):
    """A FITS file as Siril's savefits would lay it out: frame uint16 [C, H, W] (rows in fits.data order)."""
    c, h, w = frame.shape
    cards = [_card("SIMPLE", "T"), _card("BITPIX", str(bitpix)), _card("NAXIS", "3" if c == 3 else "2"),
             _card("NAXIS1", str(w)), _card("NAXIS2", str(h))]
    if c == 3:
        cards.append(_card("NAXIS3", "3"))
    cards += [_card("BZERO", "32768" if bitpix == 16 else "0"), _card("BSCALE", "1")]
    if exposure:
        cards.append(_card("EXPTIME", repr(float(exposure))))
    cards += [_card(k, v) for k, v in extra]
    cards.append("END".ljust(80))
    header = "".join(cards).encode("ascii")
    header += b" " * (-len(header) % 2880)
    if bitpix == 16:        # image_format_fits.c:715-719 TUSHORT with BZERO 32768: stored = value - 32768, big endian
        data = (frame.astype(np.int32) - 32768).astype(">i2").tobytes()
    else:                   # BYTE_IMG (:697-712 with norm != USHRT_MAX): the low byte
        data = frame.astype(np.uint8).tobytes()
    return header + data + b"\0" * (-len(data) % 2880)


def fits_decode(buf):
    """(frame uint16 [C, H, W], cards dict) of a FITS file: image_format_fits.c:112-130 (fits_read_pix TUSHORT / TBYTE)."""
    cards, off = {}, 0
    while True:
        card = buf[off:off + 80].decode("ascii")
        off += 80
        if card.startswith("END"):
            break
        if card[8] == "=":
            cards[card[:8].strip()] = card[10:].split("/")[0].strip().strip("'").strip()
    off = (off + 2879) // 2880 * 2880
    w, h = int(cards["NAXIS1"]), int(cards["NAXIS2"])
    c = int(cards.get("NAXIS3", 1)) if int(cards["NAXIS"]) == 3 else 1
    n = c * h * w
    if int(cards["BITPIX"]) == 16:
        stored = np.frombuffer(buf, dtype=">i2", count=n, offset=off).astype(np.int32)
        frame = (stored + int(float(cards.get("BZERO", 0)))).clip(0, 65535).astype(np.uint16)
    else:
        frame = np.frombuffer(buf, dtype=np.uint8, count=n, offset=off).astype(np.uint16)
    return frame.reshape(c, h, w), cards, off


def ser_encode(frames, depth_bits=16, big_endian=False, color_id=0, frame_count_in_header=True):
    """A SER file (ser.c:291-350 for the header layout): frames uint16 [N, C, H, W], rows top-down as in the file."""
    n, c, h, w = frames.shape
    hdr = bytearray(178)
    hdr[0:14] = b"LUCAM-RECORDER"
    struct.pack_into("<7i", hdr, 14, 0, color_id, 1 if big_endian else 0, w, h, depth_bits, n if frame_count_in_header else 0)
    body = bytearray()
    for f in frames:
        px = f.transpose(1, 2, 0)                           # [H, W, C]: interleaved colours
        if color_id == 101:
            px = px[:, :, ::-1]                              # BGR in the file
        if depth_bits <= 8:
            body += px.astype(np.uint8).tobytes()
        else:
            body += px.astype(">u2" if big_endian else "<u2").tobytes()
    return bytes(hdr) + bytes(body)


def ser_decode_frame(buf, index):
    """Planar uint16 [C, H, W] of frame `index`, rows as in the file: ser_read_opened_partial (ser.c:772-971) with
    ser_manage_endianess_and_depth (:476-491)."""
    color_id, big_endian, w, h, depth, count = struct.unpack_from("<6i", buf, 18)
    bps = 1 if depth <= 8 else 2
    planes = 3 if color_id in (100, 101) else 1
    size = w * h * planes * bps
    raw = buf[178 + index * size:178 + (index + 1) * size]
    if bps == 1:
        px = np.frombuffer(raw, np.uint8).astype(np.uint16)
    else:
        px = np.frombuffer(raw, ">u2" if big_endian else "<u2").astype(np.uint16)
    px = px.reshape(h, w, planes)
    if color_id == 101:
        px = px[:, :, ::-1]
    return np.ascontiguousarray(px.transpose(2, 0, 1))


def preprocess(frame, offset=None, dark=None, flat=None, level=1.0):
    """preprocess (siril.c:945-962): imoper OPER_SUB saturating at 0 (:169-173), fdiv (:252-275)."""
    g = frame.astype(np.int64)
    if offset is not None:
        g = np.maximum(g - offset, 0)
    if dark is not None:
        g = np.maximum(g - dark, 0)
    if flat is not None:
        f = np.where(flat == 0, 1, flat).astype(np.float64)
        t = np.float64(np.float32(level)) * (g.astype(np.float64) / f)
        g = np.where(t <= 0, 0, np.where(t > 65535.0, 65535, np.floor(t + 0.5))).astype(np.int64)
    return g.astype(np.uint16)


def basic_statistics(layer, nullcheck=True):
    """statistics(fit, layer, NULL, STATS_BASIC, STATS_ZERO_NULLCHECK) (src/algos/statistics.c:207-326) of one uint16
    [H, W] layer.  median: siril_stats_ushort_median (:47-64); min / max: gsl_stats_ushort_minmax over the non-zero
    pixels.  ngoodpix / mean / sigma / bgnoise come from cfitsio's fits_img_stats_ushort, which is NOT part of
    /root/reference: FnMeanSigma (mean = sum / n, sigma = sqrt(sum2 / n - mean^2) over the non-null pixels) and FnNoise1
    (per row: sigma of the differences of consecutive non-null pixels with three passes of 5-sigma clipping; the median
    over the rows that have at least two differences; times 0.70710678) are restated from its published source --
    parity unpinned."""
    h, w = layer.shape
    data = layer.ravel()
    good = data[data != 0] if nullcheck else data
    n = good.size
    if n == 0:
        return None
    g = good.astype(np.float64)
    mean = g.sum() / n
    sigma = np.sqrt((g * g).sum() / n - mean * mean) if n > 1 else 0.0
    hist = np.bincount(data, minlength=65536).astype(np.float64)
    run, median = 0.0, 0.0
    for i in range(1 if nullcheck else 0, 65536):
        run += hist[i]
        if run > n * 0.5:
            median = float(i)
            break
    sds = []
    for row in layer:
        vals = row[row != 0] if nullcheck else row
        if vals.size < 3:
            continue
        d = (vals[:-1].astype(np.int64) - vals[1:].astype(np.int64)).astype(np.float32).astype(np.float64)
        m, sd = d.mean(), np.sqrt(max((d * d).sum() / d.size - d.mean() ** 2, 0.0))
        if sd > 0.0:
            for _ in range(3):
                keep = np.abs(d - m) < 5.0 * sd
                if keep.all():
                    break
                d = d[keep]
                if d.size > 1:
                    m, sd = d.sum() / d.size, np.sqrt(max((d * d).sum() / d.size - (d.sum() / d.size) ** 2, 0.0))
                elif d.size == 1:
                    m, sd = d[0], 0.0
                else:
                    m, sd = 0.0, 0.0
        sds.append(sd)
    sds = np.sort(np.array(sds))
    noise = 0.0 if sds.size == 0 else (sds[0] if sds.size == 1 else (sds[(sds.size - 1) // 2] + sds[sds.size // 2]) / 2.0)
    return dict(total=float(h * w), ngoodpix=float(n), mean=mean, median=median, sigma=sigma, bgnoise=0.70710678 * noise,
                min=float(good.min()), max=float(good.max()))
