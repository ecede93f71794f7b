"""ctypes front end of oracle/_ref/libsiril_ref.so.  TEST INFRASTRUCTURE ONLY.

The library is the reference's src/stacking/stacking.c + src/algos/statistics.c
compiled unmodified (oracle/Makefile) and fed from memory (oracle/ref_shim/).
"""
import contextlib
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsiril_ref.so")

SUM, MEAN, MEDIAN, NORM_ONLY, ADDMAX, ADDMIN = 0, 1, 2, 3, 4, 5
# reference enums (stacking.h:14-30)
NO_REJEC, PERCENTILE, SIGMA, SIGMEDIAN, WINSORIZED, LINEARFIT = range(6)
NO_NORM, ADDITIVE, MULTIPLICATIVE, ADDITIVE_SCALING, MULTIPLICATIVE_SCALING = range(5)


class _Config(C.Structure):
    _fields_ = [
        ("frames", C.c_void_p),
        ("n_frames", C.c_int), ("width", C.c_int), ("height", C.c_int), ("channels", C.c_int),
        ("method", C.c_int), ("rejection", C.c_int), ("normalize", C.c_int),
        ("sig_low", C.c_double), ("sig_high", C.c_double),
        ("reference_image", C.c_int), ("row_order", C.c_int),
        ("shiftx", C.c_void_p), ("shifty", C.c_void_p),
        ("max_thread", C.c_int), ("max_number_of_rows", C.c_int), ("verbose", C.c_int),
    ]


_lib = None


@contextlib.contextmanager
def _quiet_stdout(enable=True):
    """The reference printf()s its block layout to stdout; keep test logs (and
    bench.py's single JSON line) clean by pointing fd 1 at /dev/null meanwhile."""
    if not enable:
        yield
        return
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        C.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.oref_run.restype = C.c_int
        _lib.oref_run.argtypes = [C.POINTER(_Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oref_frame_stats.restype = C.c_int
        _lib.oref_frame_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_long)]
        _lib.oref_num_procs.restype = C.c_int
        _lib.oref_last_hi.restype = C.c_int
    return _lib


def num_procs():
    return lib().oref_num_procs()


def run(frames, method, rejection=NO_REJEC, normalize=NO_NORM, sig=(0.0, 0.0),
        reference_image=-1, row_order=0, shiftx=None, shifty=None, max_thread=0,
        max_number_of_rows=0, want_coeff=False, verbose=False):
    """frames: uint16 [N, C, H, W].  Returns dict(out [C,H,W], irej [3,2], coeff [3,N]|None,
    retval, t_method, t_norm)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint16)
    n, c, h, w = frames.shape
    out = np.zeros((c, h, w), dtype=np.uint16)
    irej = np.zeros(6, dtype=np.uint64)
    coeff = np.zeros((3, n), dtype=np.float64) if want_coeff else None
    timing = np.zeros(2, dtype=np.float64)
    sx = None if shiftx is None else np.ascontiguousarray(shiftx, dtype=np.int32)
    sy = None if shifty is None else np.ascontiguousarray(shifty, dtype=np.int32)
    cfg = _Config(frames.ctypes.data, n, w, h, c, method, rejection, normalize,
                  float(sig[0]), float(sig[1]), reference_image, row_order,
                  None if sx is None else sx.ctypes.data, None if sy is None else sy.ctypes.data,
                  max_thread, max_number_of_rows, int(verbose))
    with _quiet_stdout(not verbose):
        rv = lib().oref_run(C.byref(cfg), out.ctypes.data, irej.ctypes.data,
                            None if coeff is None else coeff.ctypes.data, timing.ctypes.data)
    if coeff is not None and normalize == NO_NORM:
        coeff[0], coeff[1], coeff[2] = 0.0, 1.0, 1.0     # compute_normalization's defaults (stacking.c:130-134)
    return dict(out=out, irej=irej.reshape(3, 2), coeff=coeff, retval=rv, hi=lib().oref_last_hi(),
                t_method=float(timing[0]), t_norm=float(timing[1]))


def frame_stats(frame):
    """frame: uint16 [C, H, W] -> (location, scale, normValue, ngoodpix) as normalization sees them."""
    frame = np.ascontiguousarray(frame, dtype=np.uint16)
    c, h, w = frame.shape
    loc, sc, nv, ng = C.c_double(), C.c_double(), C.c_double(), C.c_long()
    rv = lib().oref_frame_stats(frame.ctypes.data, w, h, c, C.byref(loc), C.byref(sc), C.byref(nv), C.byref(ng))
    if rv:
        raise RuntimeError("statistics() returned NULL (no non-zero pixel)")
    return loc.value, sc.value, nv.value, ng.value
