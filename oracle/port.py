"""ctypes front end of oracle/libsiril_oracle.so (siril_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsiril_oracle.so")

NO_REJEC, PERCENTILE, SIGMA, SIGMEDIAN, WINSORIZED, LINEARFIT = range(6)
NO_NORM, ADDITIVE, MULTIPLICATIVE, ADDITIVE_SCALING, MULTIPLICATIVE_SCALING = range(5)
# sd_mode: SD_GSL = the reference's formulation of sigma (GSL's running mean / variance in x87 long double);
# SD_EXACT = the GPU's (exact integer sums, one rounding).  FIT_EXACT is a flag OR-ed in: the straight line
# of LINEARFIT from exact integer sums instead of gsl_fit_linear's running means.  The GPU screens with that
# form but finishes every undecided pixel with the literal recurrences, so it is compared WITHOUT the flag.
SD_GSL, SD_EXACT, FIT_EXACT = 0, 1, 2


class _Args(C.Structure):
    _fields_ = [
        ("frames", C.c_void_p),
        ("n_frames", C.c_int), ("width", C.c_int), ("height", C.c_int), ("channels", C.c_int),
        ("normalize", C.c_int), ("rejection", C.c_int),
        ("sig_low", C.c_double), ("sig_high", C.c_double),
        ("coeff", C.c_void_p), ("shiftx", C.c_void_p), ("shifty", C.c_void_p),
        ("row_order", C.c_int), ("sd_mode", C.c_int), ("threads", C.c_int),
    ]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libsiril_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.so_round_to_word.restype = C.c_uint16
        _lib.so_round_to_word.argtypes = [C.c_double]
        _lib.so_reject_column.restype = C.c_uint16
        _lib.so_reject_column.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _lib


def _args(frames, normalize, rejection, sig, coeff, shiftx, shifty, row_order, sd_mode, threads):
    frames = np.ascontiguousarray(frames, dtype=np.uint16)
    n, c, h, w = frames.shape
    keep = [frames]
    def ptr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data
    a = _Args(frames.ctypes.data, n, w, h, c, normalize, rejection, float(sig[0]), float(sig[1]),
              ptr(coeff, np.float64), ptr(shiftx, np.int32), ptr(shifty, np.int32),
              row_order, sd_mode, threads)
    return a, keep, (n, c, h, w)


def stack_sum(frames, shiftx=None, shifty=None, row_order=0):
    a, keep, (n, c, h, w) = _args(frames, 0, 0, (0, 0), None, shiftx, shifty, row_order, 0, 0)
    out = np.zeros((c, h, w), np.uint16)
    mx = C.c_uint64()
    rv = lib().so_stack_sum(C.byref(a), C.c_void_p(out.ctypes.data), C.byref(mx))
    return dict(out=out, retval=rv, maxim=mx.value)


def stack_extreme(frames, take_max, shiftx=None, shifty=None, row_order=0):
    """stack_addmax (take_max=True) / stack_addmin; 'hi' is what the reference leaves in gfit.hi."""
    a, keep, (n, c, h, w) = _args(frames, 0, 0, (0, 0), None, shiftx, shifty, row_order, 0, 0)
    out = np.zeros((c, h, w), np.uint16)
    ex = C.c_uint16()
    rv = lib().so_stack_extreme(C.byref(a), int(bool(take_max)), C.c_void_p(out.ctypes.data), C.byref(ex))
    return dict(out=out, retval=rv, hi=ex.value)


def stack_median(frames, normalize=NO_NORM, coeff=None, row_order=0, threads=0):
    a, keep, (n, c, h, w) = _args(frames, normalize, 0, (0, 0), coeff, None, None, row_order, 0, threads)
    out = np.zeros((c, h, w), np.uint16)
    rv = lib().so_stack_median(C.byref(a), C.c_void_p(out.ctypes.data))
    return dict(out=out, retval=rv)


def stack_mean(frames, rejection=NO_REJEC, sig=(0.0, 0.0), normalize=NO_NORM, coeff=None,
               shiftx=None, shifty=None, row_order=0, sd_mode=SD_GSL, threads=0):
    a, keep, (n, c, h, w) = _args(frames, normalize, rejection, sig, coeff, shiftx, shifty,
                                  row_order, sd_mode, threads)
    out = np.zeros((c, h, w), np.uint16)
    irej = np.zeros(6, np.uint64)
    g, cap = C.c_uint64(), C.c_uint64()
    gmap = np.zeros((c, h, w), np.uint8)
    rv = lib().so_stack_mean_map(C.byref(a), C.c_void_p(out.ctypes.data), C.c_void_p(irej.ctypes.data),
                                 C.byref(g), C.byref(cap), C.c_void_p(gmap.ctypes.data))
    return dict(out=out, retval=rv, irej=irej.reshape(3, 2), guard_fired=g.value, cap_hit=cap.value,
                guard_map=gmap.astype(bool))


def reject_column(stack, rejection, sig, sd_mode=SD_GSL):
    s = np.array(stack, dtype=np.uint16)
    rej = np.zeros(2, np.uint64)
    g, cap = C.c_int(0), C.c_int(0)
    v = lib().so_reject_column(s.ctypes.data, len(s), rejection, float(sig[0]), float(sig[1]), sd_mode,
                               rej.ctypes.data, C.byref(g), C.byref(cap))
    return int(v), rej, g.value, cap.value


def frame_stats_sorted(frame):
    frame = np.ascontiguousarray(frame, dtype=np.uint16)
    c, h, w = frame.shape
    loc, sc, nv, ng = C.c_double(), C.c_double(), C.c_double(), C.c_long()
    rv = lib().so_frame_stats_sorted(C.c_void_p(frame.ctypes.data), w, h, c, C.byref(loc), C.byref(sc),
                                     C.byref(nv), C.byref(ng))
    if rv:
        raise RuntimeError("no non-zero pixel")
    return loc.value, sc.value, nv.value, ng.value


def histogram(frame):
    frame = np.ascontiguousarray(frame, dtype=np.uint16)
    c, h, w = frame.shape
    hist = np.zeros(65536, np.uint32)
    mx = C.c_uint16()
    lib().so_histogram(C.c_void_p(frame.ctypes.data), w, h, c, C.c_void_p(hist.ctypes.data), C.byref(mx))
    return hist, mx.value


def ikss_hist(hist, norm_value=65535.0):
    hist = np.ascontiguousarray(hist, dtype=np.uint32)
    loc, sc, it = C.c_double(), C.c_double(), C.c_int()
    lib().so_ikss_hist(C.c_void_p(hist.ctypes.data), C.c_double(norm_value), C.byref(loc), C.byref(sc), C.byref(it))
    return loc.value, sc.value, it.value


def norm_coeffs(mode, location, scale, ref_image=0):
    location = np.ascontiguousarray(location, np.float64)
    scale = np.ascontiguousarray(scale, np.float64)
    n = len(location)
    coeff = np.zeros((3, n), np.float64)
    lib().so_norm_coeffs(mode, n, ref_image, C.c_void_p(location.ctypes.data), C.c_void_p(scale.ctypes.data),
                         C.c_void_p(coeff.ctypes.data))
    return coeff


def frame_stats_hist(frame):
    """location/scale of one [C,H,W] frame through the histogram form of IKSS."""
    hist, mx = histogram(frame)
    norm = 255.0 if mx <= 255 else 65535.0
    hist = hist.copy()
    hist[0] = 0
    loc, sc, it = ikss_hist(hist, norm)
    return loc, sc, norm, it
