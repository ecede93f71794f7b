"""Host-side Python mirror of libsiril_stack_cuda's C ABI (include/siril_stack_cuda.h).

Used by the tests and the benchmark; a C caller (Siril itself)
binds the same symbols directly.  Everything here goes through the shared
library -- there is no CPU path: importing fails loudly when the library has
not been built, and every operator raises StackError when no CUDA device is
usable.

Operator layer (reference: src/stacking/stacking.c):
    seq = Sequence(frames_u16[N, C, H, W])
    stack_summing(seq)                      # stack_summing            :196
    stack_addmax(seq) / stack_addmin(seq)   # stack_addmax :824 / stack_addmin :979
    stack_median(seq, normalize=...)        # stack_median             :362
    stack_mean_with_rejection(seq, ...)     # stack_mean_with_rejection :1189
    compute_normalization(seq, mode)        # compute_normalization    :125
Device layer: dev.* wrap ssc_dev_* on raw device pointers (torch tensors' data_ptr()).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSC_LIB", os.path.join(_HERE, "libsiril_stack_cuda.so"))	# SSC_LIB: experiments only
if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libsiril_stack_cuda.so is not built (run `python __graft_entry__.py`); "
        "this package has no CPU fallback")
_lib = C.CDLL(LIB_PATH)

# reference enums (src/stacking/stacking.h:14-30)
NO_REJEC, PERCENTILE, SIGMA, SIGMEDIAN, WINSORIZED, LINEARFIT = range(6)
NO_NORM, ADDITIVE, MULTIPLICATIVE, ADDITIVE_SCALING, MULTIPLICATIVE_SCALING = range(5)
ROWS_FITS, ROWS_SER = 0, 1
FRAMES_U16, FRAMES_FITS16, FRAMES_RAW8, FRAMES_RAW16_BE, FRAMES_SER_RGB = range(5)
HIST_STRIDE = 65540
N_COUNTERS = 8

ERR_CUDA, ERR_UNBOUND, ERR_ARG, ERR_NO_RESULT = -100, -101, -102, -103


class StackError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = _lib.ssc_last_error()
        super().__init__("%s returned %d: %s" % (where, code, msg.decode() if msg else ""))


class _Timeval(C.Structure):
    _fields_ = [("tv_sec", C.c_long), ("tv_usec", C.c_long)]


FILTER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double)


class StackingArgs(C.Structure):
    """struct stacking_args, src/stacking/stacking.h:38-56 (sizeof 224)."""
    _fields_ = [
        ("method", C.c_void_p),
        ("seq", C.c_void_p),
        ("filtering_criterion", FILTER_FN),
        ("filtering_parameter", C.c_double),
        ("nb_images_to_stack", C.c_int),
        ("image_indices", C.POINTER(C.c_int)),
        ("description", C.c_char * 100),
        ("output_filename", C.c_char_p),
        ("output_overwrite", C.c_int),
        ("t_start", _Timeval),
        ("retval", C.c_int),
        ("max_number_of_rows", C.c_int),
        ("sig", C.c_double * 2),
        ("type_of_rejection", C.c_int),
        ("normalize", C.c_int),
        ("force_norm", C.c_int),
    ]


class NormCoeff(C.Structure):
    _fields_ = [("offset", C.POINTER(C.c_double)), ("mul", C.POINTER(C.c_double)),
                ("scale", C.POINTER(C.c_double))]


class SequenceView(C.Structure):
    _fields_ = [
        ("number", C.c_int), ("nb_layers", C.c_int), ("rx", C.c_int), ("ry", C.c_int),
        ("reference_image", C.c_int), ("row_order", C.c_int),
        ("frames", C.POINTER(C.c_void_p)), ("frames_on_device", C.c_int),
        ("shiftx", C.POINTER(C.c_int)), ("shifty", C.POINTER(C.c_int)),
        ("exposure", C.POINTER(C.c_double)),
        ("frame_format", C.c_int), ("ser_bytes_per_sample", C.c_int), ("ser_big_endian", C.c_int), ("ser_bgr", C.c_int),
        ("pp_offset", C.c_void_p), ("pp_dark", C.c_void_p), ("pp_flat", C.c_void_p), ("pp_flat_level", C.c_float),
    ]


class FitsInfo(C.Structure):
    _fields_ = [("bitpix", C.c_int), ("naxis", C.c_int), ("naxes", C.c_long * 3), ("bzero", C.c_double), ("bscale", C.c_double),
                ("exposure", C.c_double), ("hi", C.c_ushort), ("lo", C.c_ushort), ("data_offset", C.c_size_t),
                ("data_bytes", C.c_size_t), ("frame_format", C.c_int)]


class SerInfo(C.Structure):
    _fields_ = [("color_id", C.c_int), ("big_endian", C.c_int), ("width", C.c_int), ("height", C.c_int), ("bit_depth", C.c_int),
                ("frame_count", C.c_int), ("bytes_per_sample", C.c_int), ("planes", C.c_int), ("frame_bytes", C.c_size_t),
                ("data_offset", C.c_size_t), ("frame_format", C.c_int)]


class Result(C.Structure):
    _fields_ = [
        ("data", C.c_void_p), ("pdata", C.c_void_p * 3),
        ("rx", C.c_uint), ("ry", C.c_uint), ("naxis", C.c_int), ("naxes", C.c_long * 3),
        ("exposure", C.c_double), ("hi", C.c_ushort),
        ("irej", (C.c_uint64 * 2) * 3), ("guard_fired", C.c_uint64), ("cap_hit", C.c_uint64),
        ("norm_offset0", C.c_double),
        ("t_h2d_ms", C.c_double), ("t_device_ms", C.c_double), ("t_d2h_ms", C.c_double),
        ("t_total_ms", C.c_double),
    ]


class Band(C.Structure):
    """ssc_band: a row band of all frames resident in device memory."""
    _fields_ = [
        ("frames", C.c_void_p), ("frame_stride", C.c_size_t), ("layer_stride", C.c_size_t),
        ("n_frames", C.c_int), ("layers", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("buf_row0", C.c_int), ("buf_rows", C.c_int), ("out_row0", C.c_int), ("out_rows", C.c_int),
        ("row_order", C.c_int), ("shiftx", C.c_void_p), ("shifty", C.c_void_p),
    ]


LOG_FN = C.CFUNCTYPE(None, C.c_char_p)
PROGRESS_FN = C.CFUNCTYPE(None, C.c_char_p, C.c_double)
CONTINUE_FN = C.CFUNCTYPE(C.c_int)


class Callbacks(C.Structure):
    _fields_ = [("log", LOG_FN), ("progress", PROGRESS_FN), ("should_continue", CONTINUE_FN)]


# every symbol include/siril_stack_cuda.h declares
EXPORTS = [
    "ssc_bind_sequence", "ssc_unbind_sequence", "ssc_set_callbacks", "ssc_take_result", "ssc_release_workspace",
    "ssc_stack_summing", "ssc_stack_median", "ssc_stack_mean_with_rejection", "ssc_stack_addmax", "ssc_stack_addmin",
    "ssc_compute_normalization", "ssc_set_devices", "ssc_device_count", "ssc_set_kernel_path", "ssc_last_error",
    "ssc_version", "ssc_dev_stack_median", "ssc_dev_stack_mean", "ssc_dev_sum_accumulate",
    "ssc_dev_sum_finish", "ssc_dev_stack_extreme", "ssc_dev_histogram", "ssc_dev_ikss", "ssc_dev_norm_coeffs",
    "ssc_dev_synth_frames",
    "ssc_peer_create", "ssc_peer_export", "ssc_peer_import", "ssc_peer_attach_local", "ssc_peer_connect",
    "ssc_peer_destroy", "ssc_peer_world", "ssc_peer_rank", "ssc_peer_status", "ssc_peer_last_error",
    "ssc_dev_histogram_scatter", "ssc_dev_ikss_gather", "ssc_dev_norm_coeffs_gathered",
    "ssc_dev_counters_combine", "ssc_dev_max_combine",
    "ssc_fits_parse_header", "ssc_ser_parse_header", "ssc_savefits", "ssc_io_last_error",
    "ssc_dev_decode_fits16", "ssc_dev_decode_ser", "ssc_dev_preprocess", "ssc_image_statistics",
]

_lib.ssc_last_error.restype = C.c_char_p
_lib.ssc_version.restype = C.c_char_p
_lib.ssc_bind_sequence.argtypes = [C.c_void_p, C.POINTER(SequenceView)]
_lib.ssc_unbind_sequence.argtypes = [C.c_void_p]
_lib.ssc_set_callbacks.argtypes = [C.POINTER(Callbacks)]
_lib.ssc_set_callbacks.restype = None
_lib.ssc_take_result.argtypes = [C.POINTER(Result)]
for _n in ("ssc_stack_summing", "ssc_stack_median", "ssc_stack_mean_with_rejection", "ssc_stack_addmax",
           "ssc_stack_addmin"):
    getattr(_lib, _n).argtypes = [C.POINTER(StackingArgs)]
_lib.ssc_compute_normalization.argtypes = [C.POINTER(StackingArgs), C.POINTER(NormCoeff), C.c_int]
_lib.ssc_set_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
_lib.ssc_dev_stack_median.argtypes = [C.POINTER(Band), C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_int, C.c_void_p]
_lib.ssc_dev_stack_mean.argtypes = [C.POINTER(Band), C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                    C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
_lib.ssc_dev_sum_accumulate.argtypes = [C.POINTER(Band), C.c_void_p, C.c_void_p, C.c_void_p]
_lib.ssc_dev_sum_finish.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
_lib.ssc_dev_stack_extreme.argtypes = [C.POINTER(Band), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
_lib.ssc_dev_histogram.argtypes = [C.POINTER(Band), C.c_void_p, C.c_void_p]
_lib.ssc_dev_ikss.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
_lib.ssc_dev_norm_coeffs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
_lib.ssc_dev_synth_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_uint64, C.c_void_p]
_lib.ssc_peer_last_error.restype = C.c_char_p
_lib.ssc_peer_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
_lib.ssc_peer_export.argtypes = [C.c_void_p, C.c_void_p]
_lib.ssc_peer_import.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
_lib.ssc_peer_attach_local.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
_lib.ssc_peer_connect.argtypes = [C.c_void_p]
_lib.ssc_peer_destroy.argtypes = [C.c_void_p]
_lib.ssc_peer_status.argtypes = [C.c_void_p, C.c_void_p]
_lib.ssc_dev_histogram_scatter.argtypes = [C.POINTER(Band), C.c_void_p, C.c_int, C.c_void_p]
_lib.ssc_dev_ikss_gather.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
_lib.ssc_dev_norm_coeffs_gathered.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]
_lib.ssc_dev_counters_combine.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
_lib.ssc_dev_max_combine.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
_lib.ssc_io_last_error.restype = C.c_char_p
_lib.ssc_fits_parse_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(FitsInfo)]
_lib.ssc_ser_parse_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(SerInfo)]
_lib.ssc_savefits.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ushort, C.c_ushort, C.c_double, C.c_int]
_lib.ssc_dev_decode_fits16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
_lib.ssc_dev_decode_ser.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
_lib.ssc_image_statistics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
_lib.ssc_dev_preprocess.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def lib():
    return _lib


def version():
    return _lib.ssc_version().decode()


def device_count():
    return _lib.ssc_device_count()


def set_kernel_path(mode):
    """0 = automatic kernel choice, 1 = general kernel only, 2 = automatic with the exact normalization only (diagnostic)."""
    rc = _lib.ssc_set_kernel_path(int(mode))
    if rc:
        raise StackError(rc, "ssc_set_kernel_path")


def release_workspace():
    _lib.ssc_release_workspace()


def set_devices(ids):
    arr = (C.c_int * len(ids))(*ids)
    rc = _lib.ssc_set_devices(arr, len(ids))
    if rc:
        raise StackError(rc, "ssc_set_devices")


_cb_keepalive = None


def set_callbacks(log=None, progress=None, should_continue=None):
    """Bind the host hooks (siril_log_message / set_progress_bar_data / get_thread_run)."""
    global _cb_keepalive
    cb = Callbacks(
        LOG_FN(lambda m: log(m.decode())) if log else LOG_FN(),
        PROGRESS_FN(lambda t, f: progress(t.decode() if t else None, f)) if progress else PROGRESS_FN(),
        CONTINUE_FN(lambda: int(bool(should_continue()))) if should_continue else CONTINUE_FN())
    _cb_keepalive = cb
    _lib.ssc_set_callbacks(C.byref(cb))


class Sequence:
    """An in-memory sequence bound to the library: what `sequence *` + its files are to Siril.

    frames: uint16 array [number, layers, ry, rx] (host; pinned memory uploads fastest) or a list
    of per-frame [layers, ry, rx] arrays / raw pointers.  shiftx/shifty: regparam[reglayer][i]."""

    _next_handle = 0x1000

    def __init__(self, frames, reference_image=-1, row_order=ROWS_FITS, shiftx=None, shifty=None,
                 exposure=None, on_device=False, shape=None, frame_format=FRAMES_U16, ser_bytes_per_sample=0,
                 ser_big_endian=0, ser_bgr=0, pp_offset=None, pp_dark=None, pp_flat=None, pp_flat_level=1.0):
        """frame_format != FRAMES_U16: frames is a list of raw pointers (ints) to each frame's first sample as it lies
        in the file (FITS data unit / SER frame), shape = (number, layers, ry, rx); pp_*: preprocess masters, uint16
        [layers, ry, rx]."""
        if on_device or frame_format != FRAMES_U16 or shape is not None:
            self._keep = frames
            number, layers, ry, rx = shape
            ptrs = [int(p) for p in frames]
        else:
            if isinstance(frames, np.ndarray):
                assert frames.dtype == np.uint16 and frames.ndim == 4
                frames = np.ascontiguousarray(frames)
                number, layers, ry, rx = frames.shape
                self._keep = frames
                ptrs = [frames[i].ctypes.data for i in range(number)]
            else:
                frames = [np.ascontiguousarray(f, dtype=np.uint16) for f in frames]
                number = len(frames)
                layers, ry, rx = frames[0].shape
                self._keep = frames
                ptrs = [f.ctypes.data for f in frames]
        self.number, self.layers, self.ry, self.rx = number, layers, ry, rx
        self._ptrs = (C.c_void_p * number)(*ptrs)
        self._sx = None if shiftx is None else (C.c_int * number)(*[int(v) for v in shiftx])
        self._sy = None if shifty is None else (C.c_int * number)(*[int(v) for v in shifty])
        self._exp = None if exposure is None else (C.c_double * number)(*[float(v) for v in exposure])
        Sequence._next_handle += 16
        self.handle = Sequence._next_handle	# stands in for the `sequence *` key
        self._masters = [None if m is None else np.ascontiguousarray(m, dtype=np.uint16) for m in (pp_offset, pp_dark, pp_flat)]
        mp = [None if m is None else m.ctypes.data for m in self._masters]
        view = SequenceView(number, layers, rx, ry, reference_image, row_order,
                            C.cast(self._ptrs, C.POINTER(C.c_void_p)), int(on_device),
                            self._sx, self._sy, self._exp, frame_format, ser_bytes_per_sample, ser_big_endian, ser_bgr,
                            mp[0], mp[1], mp[2], float(pp_flat_level))
        rc = _lib.ssc_bind_sequence(self.handle, C.byref(view))
        if rc:
            raise StackError(rc, "ssc_bind_sequence")

    def close(self):
        if self.handle is not None:
            _lib.ssc_unbind_sequence(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StackResult:
    def __init__(self, r, coeff=None):
        c, h, w = int(r.naxes[2]), int(r.ry), int(r.rx)
        n = c * h * w
        self._ptr = r.data			# ownership was passed to us, as with gfit.data: freed in __del__
        buf = (C.c_uint16 * n).from_address(r.data)
        self.data = np.frombuffer(buf, dtype=np.uint16).reshape(c, h, w)
        self.exposure = r.exposure
        self.hi = int(r.hi)
        self.irej = np.array([[r.irej[i][0], r.irej[i][1]] for i in range(3)], dtype=np.uint64)
        self.guard_fired = int(r.guard_fired)
        self.cap_hit = int(r.cap_hit)
        self.t_h2d_ms, self.t_device_ms = r.t_h2d_ms, r.t_device_ms
        self.t_d2h_ms, self.t_total_ms = r.t_d2h_ms, r.t_total_ms
        self.coeff = coeff

    def __del__(self):
        if getattr(self, "_ptr", None):
            self.data = None
            _libc.free(self._ptr)
            self._ptr = None


def _filter_all(seq, nb_img, any_):
    return 1


_FILTER_ALL = FILTER_FN(_filter_all)


def make_args(seq, image_indices=None, normalize=NO_NORM, rejection=NO_REJEC, sig=(0.0, 0.0),
              filtering_criterion=None):
    if image_indices is None:
        image_indices = list(range(seq.number))
    idx = (C.c_int * len(image_indices))(*image_indices)
    args = StackingArgs()
    args.seq = seq.handle
    args.filtering_criterion = filtering_criterion if filtering_criterion is not None else _FILTER_ALL
    args.filtering_parameter = 0.0
    args.nb_images_to_stack = len(image_indices)
    args.image_indices = C.cast(idx, C.POINTER(C.c_int))
    args._idx_keep = idx
    args.retval = -1
    args.max_number_of_rows = seq.ry
    args.sig[0], args.sig[1] = float(sig[0]), float(sig[1])
    args.type_of_rejection = rejection
    args.normalize = normalize
    args.force_norm = 1
    return args


def _take(coeff=None):
    r = Result()
    rc = _lib.ssc_take_result(C.byref(r))
    if rc:
        raise StackError(rc, "ssc_take_result")
    return StackResult(r, coeff)


def compute_normalization(seq, mode, image_indices=None):
    """compute_normalization (stacking.c:125) -> float64 [3, N] = offset, mul, scale."""
    args = make_args(seq, image_indices, normalize=mode)
    n = args.nb_images_to_stack
    off, mul, sc = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
    nc = NormCoeff(off, mul, sc)
    rc = _lib.ssc_compute_normalization(C.byref(args), C.byref(nc), mode)
    if rc:
        raise StackError(rc, "ssc_compute_normalization")
    return np.array([list(off), list(mul), list(sc)], dtype=np.float64)


def stack_summing(seq, image_indices=None, filtering_criterion=None):
    args = make_args(seq, image_indices, filtering_criterion=filtering_criterion)
    rc = _lib.ssc_stack_summing(C.byref(args))
    if rc:
        raise StackError(rc, "ssc_stack_summing")
    return _take()


def stack_addmax(seq, image_indices=None, filtering_criterion=None):
    args = make_args(seq, image_indices, filtering_criterion=filtering_criterion)
    rc = _lib.ssc_stack_addmax(C.byref(args))
    if rc:
        raise StackError(rc, "ssc_stack_addmax")
    return _take()


def stack_addmin(seq, image_indices=None, filtering_criterion=None):
    args = make_args(seq, image_indices, filtering_criterion=filtering_criterion)
    rc = _lib.ssc_stack_addmin(C.byref(args))
    if rc:
        raise StackError(rc, "ssc_stack_addmin")
    return _take()


def stack_median(seq, normalize=NO_NORM, image_indices=None, want_coeff=False):
    coeff = compute_normalization(seq, normalize, image_indices) if (want_coeff and normalize) else None
    args = make_args(seq, image_indices, normalize=normalize)
    rc = _lib.ssc_stack_median(C.byref(args))
    if rc:
        raise StackError(rc, "ssc_stack_median")
    return _take(coeff)


def stack_mean_with_rejection(seq, rejection=NO_REJEC, sig=(0.0, 0.0), normalize=NO_NORM,
                              image_indices=None, want_coeff=True):
    coeff = compute_normalization(seq, normalize, image_indices) if (want_coeff and normalize) else None
    args = make_args(seq, image_indices, normalize=normalize, rejection=rejection, sig=sig)
    rc = _lib.ssc_stack_mean_with_rejection(C.byref(args))
    if rc:
        raise StackError(rc, "ssc_stack_mean_with_rejection")
    return _take(coeff)


class dev:
    """Device layer: ssc_dev_* on raw device pointers (ints) and a raw cudaStream_t (int)."""

    @staticmethod
    def band(frames_ptr, n_frames, layers, height, width, out_row0=0, out_rows=None, buf_row0=0,
             buf_rows=None, row_order=ROWS_FITS, shiftx_ptr=None, shifty_ptr=None,
             frame_stride=None, layer_stride=None):
        buf_rows = height if buf_rows is None else buf_rows
        out_rows = height - out_row0 if out_rows is None else out_rows
        layer_stride = buf_rows * width if layer_stride is None else layer_stride
        frame_stride = layer_stride * layers if frame_stride is None else frame_stride
        return Band(frames_ptr, frame_stride, layer_stride, n_frames, layers, width, height,
                    buf_row0, buf_rows, out_row0, out_rows, row_order, shiftx_ptr, shifty_ptr)

    @staticmethod
    def _check(rc, where):
        if rc:
            raise StackError(rc, where)

    @staticmethod
    def stack_median(band, out_ptr, normalize=NO_NORM, coeff_ptr=None, out_layer_stride=None,
                     out_buf_row0=None, stream=0):
        ols = band.out_rows * band.width if out_layer_stride is None else out_layer_stride
        ob0 = band.out_row0 if out_buf_row0 is None else out_buf_row0
        dev._check(_lib.ssc_dev_stack_median(C.byref(band), normalize, coeff_ptr, out_ptr, ols, ob0, stream),
                   "ssc_dev_stack_median")

    @staticmethod
    def stack_mean(band, out_ptr, counters_ptr, rejection=NO_REJEC, sig=(0.0, 0.0), normalize=NO_NORM,
                   coeff_ptr=None, out_layer_stride=None, out_buf_row0=None, stream=0):
        ols = band.out_rows * band.width if out_layer_stride is None else out_layer_stride
        ob0 = band.out_row0 if out_buf_row0 is None else out_buf_row0
        dev._check(_lib.ssc_dev_stack_mean(C.byref(band), normalize, coeff_ptr, rejection, float(sig[0]),
                                           float(sig[1]), out_ptr, ols, ob0, counters_ptr, stream),
                   "ssc_dev_stack_mean")

    @staticmethod
    def sum_accumulate(band, acc_ptr, maxim_ptr, stream=0):
        dev._check(_lib.ssc_dev_sum_accumulate(C.byref(band), acc_ptr, maxim_ptr, stream), "ssc_dev_sum_accumulate")

    @staticmethod
    def stack_extreme(band, take_max, out_ptr, extreme_ptr, stream=0):
        dev._check(_lib.ssc_dev_stack_extreme(C.byref(band), int(bool(take_max)), out_ptr, extreme_ptr, stream),
                   "ssc_dev_stack_extreme")

    @staticmethod
    def sum_finish(acc_ptr, count, maxim_ptr, out_ptr, stream=0):
        dev._check(_lib.ssc_dev_sum_finish(acc_ptr, count, maxim_ptr, out_ptr, stream), "ssc_dev_sum_finish")

    @staticmethod
    def histogram(band, hist_ptr, stream=0):
        dev._check(_lib.ssc_dev_histogram(C.byref(band), hist_ptr, stream), "ssc_dev_histogram")

    @staticmethod
    def ikss(hist_ptr, n_frames, stats_ptr, scratch_ptr, stream=0):
        dev._check(_lib.ssc_dev_ikss(hist_ptr, n_frames, stats_ptr, scratch_ptr, stream), "ssc_dev_ikss")

    @staticmethod
    def norm_coeffs(stats_ptr, n_frames, ref_image, mode, coeff_ptr, status_ptr=None, stream=0):
        dev._check(_lib.ssc_dev_norm_coeffs(stats_ptr, n_frames, ref_image, mode, coeff_ptr, status_ptr, stream),
                   "ssc_dev_norm_coeffs")

    @staticmethod
    def synth_frames(frames_ptr, n_frames, layers, rows, width, first_frame=0, row0=0, full_height=None,
                     kind=0, seed=0x5349524C, stream=0):
        full_height = rows if full_height is None else full_height
        dev._check(_lib.ssc_dev_synth_frames(frames_ptr, n_frames, layers, rows, width, first_frame, row0,
                                             full_height, kind, seed, stream), "ssc_dev_synth_frames")


PEER_HANDLE_BYTES = 64
PEER_TIMEOUT = -110


class PeerError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = _lib.ssc_peer_last_error()
        super().__init__("%s returned %d: %s" % (where, code, msg.decode() if msg else ""))


class PeerGroup:
    """ssc_peer_group: this GPU's end of the NVLink exchanges (include/siril_stack_cuda.h).  Created on the
    current CUDA device.  One process per GPU: export() / import_() the 64-byte handles (the host moves them,
    e.g. with torch.distributed.all_gather); several GPUs in one process: attach_local()."""

    def __init__(self, world, rank, n_frames_max):
        h = C.c_void_p()
        rc = _lib.ssc_peer_create(world, rank, n_frames_max, C.byref(h))
        if rc:
            raise PeerError(rc, "ssc_peer_create")
        self.h, self.world, self.rank = h, world, rank

    @staticmethod
    def _check(rc, where):
        if rc:
            raise PeerError(rc, where)

    def export(self):
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._check(_lib.ssc_peer_export(self.h, buf), "ssc_peer_export")
        return buf.raw

    def import_(self, peer_rank, handle):
        self._check(_lib.ssc_peer_import(self.h, peer_rank, C.create_string_buffer(bytes(handle), PEER_HANDLE_BYTES)),
                    "ssc_peer_import")

    def attach_local(self, peer_rank, other):
        self._check(_lib.ssc_peer_attach_local(self.h, peer_rank, other.h), "ssc_peer_attach_local")

    def connect(self):
        self._check(_lib.ssc_peer_connect(self.h), "ssc_peer_connect")

    def histogram_scatter(self, band, last_piece=True, stream=0):
        self._check(_lib.ssc_dev_histogram_scatter(C.byref(band), self.h, int(last_piece), stream), "ssc_dev_histogram_scatter")

    def ikss_gather(self, n_frames, scratch_ptr, stream=0):
        self._check(_lib.ssc_dev_ikss_gather(self.h, n_frames, scratch_ptr, stream), "ssc_dev_ikss_gather")

    def norm_coeffs_gathered(self, n_frames, ref_image, mode, coeff_ptr, status_ptr=None, stats_ptr=None, stream=0):
        self._check(_lib.ssc_dev_norm_coeffs_gathered(self.h, n_frames, ref_image, mode, coeff_ptr, status_ptr, stats_ptr,
                                                      stream), "ssc_dev_norm_coeffs_gathered")

    def counters_combine(self, local_ptr, total_ptr, stream=0):
        self._check(_lib.ssc_dev_counters_combine(self.h, local_ptr, total_ptr, stream), "ssc_dev_counters_combine")

    def max_combine(self, maxim_ptr, stream=0):
        self._check(_lib.ssc_dev_max_combine(self.h, maxim_ptr, stream), "ssc_dev_max_combine")

    def status(self, stream=0):
        return _lib.ssc_peer_status(self.h, stream)

    def close(self):
        if self.h:
            _lib.ssc_peer_destroy(self.h)
            self.h = None


class ImStats(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("total", "ngoodpix", "mean", "median", "sigma", "bgnoise", "min", "max", "normValue")]


def image_statistics(data, nullcheck=True):
    """ssc_image_statistics: STATS_BASIC of every layer of uint16 [layers, ry, rx] -> list of dicts."""
    data = np.ascontiguousarray(data, dtype=np.uint16)
    layers, ry, rx = data.shape
    st = (ImStats * 3)()
    rc = _lib.ssc_image_statistics(data.ctypes.data, rx, ry, layers, int(nullcheck), st)
    if rc:
        raise IoError(rc, "ssc_image_statistics")
    return [{k: getattr(st[c], k) for k, _ in ImStats._fields_} for c in range(layers)]


class IoError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = _lib.ssc_io_last_error()
        super().__init__("%s returned %d: %s" % (where, code, msg.decode() if msg else ""))


def fits_parse_header(buf):
    """ssc_fits_parse_header on a bytes-like object holding (at least) the header of a FITS file."""
    raw = (C.c_char * len(buf)).from_buffer_copy(buf)
    info = FitsInfo()
    rc = _lib.ssc_fits_parse_header(raw, len(buf), C.byref(info))
    if rc:
        raise IoError(rc, "ssc_fits_parse_header")
    return info


def ser_parse_header(buf, file_size=None):
    raw = (C.c_char * len(buf)).from_buffer_copy(buf)
    info = SerInfo()
    rc = _lib.ssc_ser_parse_header(raw, len(buf) if file_size is None else file_size, C.byref(info))
    if rc:
        raise IoError(rc, "ssc_ser_parse_header")
    return info


def savefits(path, data, hi=0, lo=0, exposure=0.0, overwrite=True):
    """ssc_savefits: data uint16 [layers, ry, rx] (a StackResult's .data)."""
    data = np.ascontiguousarray(data, dtype=np.uint16)
    layers, ry, rx = data.shape
    rc = _lib.ssc_savefits(os.fsencode(path), data.ctypes.data, rx, ry, layers, hi, lo, float(exposure), int(overwrite))
    if rc:
        raise IoError(rc, "ssc_savefits")
