/* siril_stack_cuda.h -- C ABI of libsiril_stack_cuda: Siril 0.9.7's per-pixel
 * multi-frame stacking hot path on NVIDIA B200 (sm_100a).
 *
 * Two layers, both plain C (pointers and sizes, no C++/torch types):
 *
 *  1. OPERATOR layer -- drop-in replacements for the entries of the reference's
 *     stacking_methods[] table (src/stacking/stacking.c:54-56), same
 *     `stack_method` signature, same `struct stacking_args`, same enums, same
 *     return codes.  What Siril reaches through globals (the sequence's files,
 *     the result image `gfit`, log / progress / cancel) is bound explicitly:
 *     ssc_bind_sequence(), ssc_take_result(), ssc_set_callbacks().
 *
 *  2. DEVICE layer (ssc_dev_*) -- the kernels themselves on caller-owned device
 *     memory and a caller-owned CUDA stream: what the operator layer is built
 *     from, what a multi-process (one rank per GPU) host drives between its
 *     collectives, and what the tests and the benchmark time.
 *
 * There is no CPU fallback: every entry point returns SSC_ERR_CUDA when no
 * usable device is present.
 */
#ifndef SIRIL_STACK_CUDA_H
#define SIRIL_STACK_CUDA_H

#include <stddef.h>
#include <stdint.h>
#include <sys/time.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* Types kept from the reference (skipped when Siril's own headers are in).  */
/* ------------------------------------------------------------------------- */
#ifndef STACKING_H_		/* include guard of src/stacking/stacking.h */

typedef unsigned short WORD;			/* src/core/siril.h:44 */
typedef struct sequ sequence;			/* src/core/siril.h:192,328-374 (opaque here) */
typedef int (*seq_image_filter)(sequence *seq, int nb_img, double param);	/* src/core/processing.h:5 */

struct stacking_args;
typedef int (*stack_method)(struct stacking_args *args);	/* src/stacking/stacking.h:8 */

typedef enum {					/* src/stacking/stacking.h:14-21 */
	NO_REJEC,
	PERCENTILE,
	SIGMA,
	SIGMEDIAN,
	WINSORIZED,
	LINEARFIT,
} rejection;

typedef enum {					/* src/stacking/stacking.h:24-30 */
	NO_NORM,
	ADDITIVE,
	MULTIPLICATIVE,
	ADDITIVE_SCALING,
	MULTIPLICATIVE_SCALING,
} normalization;

typedef struct normalization_coeff {		/* src/stacking/stacking.h:32-36 */
	double *offset;
	double *mul;
	double *scale;
} norm_coeff;

/* src/stacking/stacking.h:38-56 -- field order and LP64 layout kept
 * (sizeof 224; gboolean is int).  tests/test_abi.py checks the offsets against
 * the reference's own header. */
struct stacking_args {
	stack_method method;
	sequence *seq;
	seq_image_filter filtering_criterion;
	double filtering_parameter;
	int nb_images_to_stack;
	int *image_indices;
	char description[100];
	const char *output_filename;
	int output_overwrite;
	struct timeval t_start;
	int retval;
	int max_number_of_rows;
	double sig[2];			/* low and high rejection parameter */
	rejection type_of_rejection;
	normalization normalize;
	int force_norm;
};

#endif /* STACKING_H_ */

/* ------------------------------------------------------------------------- */
/* Error codes                                                               */
/* ------------------------------------------------------------------------- */
/* 0, -1, -2, -3 are the reference's own conventions:
 *  -1 cancelled / unsupported input / fewer than 2 frames / allocation or read failure
 *  -2 size mismatch or uninitialised sequence
 *  -3 (sum) frame read failure or dimension change                           */
#define SSC_OK			0
#define SSC_ERR_GENERIC		(-1)
#define SSC_ERR_SIZE		(-2)
#define SSC_ERR_READ		(-3)
/* additions of this library, outside the reference's range */
#define SSC_ERR_CUDA		(-100)	/* no device / CUDA runtime error (see ssc_last_error) */
#define SSC_ERR_UNBOUND		(-101)	/* args->seq was never given to ssc_bind_sequence */
#define SSC_ERR_ARG		(-102)	/* invalid argument */
#define SSC_ERR_NO_RESULT	(-103)	/* ssc_take_result without a finished stack */

/* ------------------------------------------------------------------------- */
/* OPERATOR layer                                                            */
/* ------------------------------------------------------------------------- */

enum {
	SSC_ROWS_FITS = 0,	/* read_opened_fits_partial convention (image_format_fits.c:599-632):
				   storage row r of the inputs -> storage row r of the result */
	SSC_ROWS_SER = 1	/* ser_read_opened_partial / ser_read_frame convention (ser.c:767,797-799):
				   the result is vertically flipped with respect to the input storage */
};

/* What the operators read from `sequence` and its files:
 * replaces seq_read_frame (sequence.c:519), seq_opened_read_region (:690) and
 * regparam[reglayer][i] by plain memory. */
typedef struct ssc_sequence_view {
	int number;			/* seq->number: frames in the sequence */
	int nb_layers;			/* 1 or 3 */
	int rx, ry;			/* width, height */
	int reference_image;		/* seq->reference_image, -1 = unset (then 0, stacking.c:143-145) */
	int row_order;			/* SSC_ROWS_* */
	const WORD *const *frames;	/* [number] pointers to planar WORD [nb_layers][ry][rx], indexed by
					   sequence image index; host memory (pinned is fastest) or device memory */
	int frames_on_device;		/* non-zero: frames[] are device pointers on the current device */
	const int *shiftx, *shifty;	/* regparam[get_registration_layer()][i].shiftx/.shifty indexed by
					   sequence image index; both NULL = no registration data */
	const double *exposure;		/* per-image exposure, may be NULL (result exposure = 0) */
	/* --- the frames as they lie in the files: decoded on the device behind the copy --- */
	int frame_format;		/* SSC_FRAMES_*: how the bytes behind frames[] are read (0 = native WORDs);
					   frames[i] points at the first sample (the FITS data unit / the SER frame),
					   e.g. into the mmap()ed file */
	int ser_bytes_per_sample;	/* SSC_FRAMES_SER_RGB: 1 or 2 (ser.c:327-329) */
	int ser_big_endian;		/* SSC_FRAMES_SER_RGB, 16 bit: the header's flag as Siril reads it (ser.h:31-42) */
	int ser_bgr;			/* SSC_FRAMES_SER_RGB: colour id 101, planes in reverse order (ser.c:954-957) */
	/* --- preprocess (src/core/siril.c:945-962) fused into the load: master frames as planar
	 * WORD [nb_layers][ry][rx] in host memory, rows in the storage order of the decoded frames; NULL = not used --- */
	const WORD *pp_offset;		/* USE_OFFSET: g = max(g - offset, 0) */
	const WORD *pp_dark;		/* USE_DARK without optimisation: g = max(g - dark, 0) */
	const WORD *pp_flat;		/* USE_FLAT: g = round_to_WORD(level * (g / max(flat, 1))) (fdiv, siril.c:252-275) */
	float pp_flat_level;		/* fdiv's coef */
} ssc_sequence_view;

enum {
	SSC_FRAMES_U16 = 0,		/* planar WORD [nb_layers][ry][rx] (also SER 16-bit mono, little endian) */
	SSC_FRAMES_FITS16 = 1,		/* FITS data unit, BITPIX 16 + BZERO 32768: big-endian signed samples, planar */
	SSC_FRAMES_RAW8 = 2,		/* planar 8-bit samples: FITS BITPIX 8 data unit, SER 8-bit mono */
	SSC_FRAMES_RAW16_BE = 3,	/* planar big-endian unsigned 16-bit: SER 16-bit mono with the big-endian flag */
	SSC_FRAMES_SER_RGB = 4		/* SER colour frames (id 100 / 101): interleaved samples, ser_* say how */
};

/* The view is copied; the frames it points at must stay valid while operators run. */
int ssc_bind_sequence(const sequence *seq, const ssc_sequence_view *view);
int ssc_unbind_sequence(const sequence *seq);

/* Host hooks replacing siril_log_message (gui/callbacks.c:2278), set_progress_bar_data
 * (:2393) and get_thread_run (core/processing.c:294).  Any member may be NULL. */
typedef struct ssc_callbacks {
	void (*log)(const char *message);
	void (*progress)(const char *text, double fraction);
	int (*should_continue)(void);	/* 0 = cancel: the operator returns -1 */
} ssc_callbacks;
void ssc_set_callbacks(const ssc_callbacks *cb);

/* What the reference leaves in the global `fits gfit` (stacking.c:323-341,776-785,
 * 1810-1827) plus the rejection counters it only logs (:1193,1811-1817). */
typedef struct ssc_result {
	WORD *data;			/* planar [naxes[2]][ry][rx], malloc()ed; the caller free()s it (as gfit.data) */
	WORD *pdata[3];
	unsigned int rx, ry;
	int naxis;
	long naxes[3];
	double exposure;		/* sum of the stacked frames' exposures */
	WORD hi;			/* sum stacking: round_to_WORD(max sum), stacking.c:324 */
	uint64_t irej[3][2];		/* per channel low / high rejections */
	uint64_t guard_fired;		/* pixels where the reference's "N - r <= 4" break would be taken
					   (history dependent in the reference) */
	uint64_t cap_hit;		/* pixels whose unbounded do-while was cut at SSC_ITER_CAP */
	double norm_offset0;		/* unused, reserved */
	double t_h2d_ms, t_device_ms, t_d2h_ms, t_total_ms;	/* CUDA-event timings of the last call */
} ssc_result;
#define SSC_ITER_CAP 4096

/* Moves the last successful result out of the library (ownership of data passes). */
int ssc_take_result(ssc_result *out);

/* stack_method-compatible operators (stacking.c:196, :362, :1189) */
int ssc_stack_summing(struct stacking_args *args);
int ssc_stack_median(struct stacking_args *args);
int ssc_stack_mean_with_rejection(struct stacking_args *args);
/* the other two entries of stacking_methods[] (stacking.c:54-56): stack_addmax (:824) and
 * stack_addmin (:979); frames chosen by args->filtering_criterion like stack_summing,
 * ssc_result.hi = the brightest / darkest result value as the reference leaves it in gfit.hi */
int ssc_stack_addmax(struct stacking_args *args);
int ssc_stack_addmin(struct stacking_args *args);
/* compute_normalization (stacking.c:125): fills coeff->offset/mul/scale[nb_images_to_stack] */
int ssc_compute_normalization(struct stacking_args *args, norm_coeff *coeff, normalization mode);

/* The operator layer keeps its device buffers (and a pinned bounce buffer for the result)
 * between calls; this returns them to the driver. */
int ssc_release_workspace(void);

/* Device selection for the operator layer: the operators split the image in row
 * bands over these devices (default: device 0 only). */
int ssc_set_devices(const int *device_ids, int count);
int ssc_device_count(void);
/* Diagnostic: 0 = pick the kernel by shape (default), 1 = always use the general
 * shared-memory kernel, 2 = pick by shape but never use the one-FMA form of the
 * normalization (which is otherwise used on launches whose coefficients it was verified
 * for, sample value by sample value).  All give identical results (tests/test_gpu_parity.py). */
int ssc_set_kernel_path(int mode);
const char *ssc_last_error(void);
const char *ssc_version(void);

/* ------------------------------------------------------------------------- */
/* DEVICE layer                                                              */
/* ------------------------------------------------------------------------- */
/* All pointers are device pointers on the current CUDA device unless stated;
 * `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 * asynchronous on that stream and return SSC_OK or SSC_ERR_CUDA / SSC_ERR_ARG.
 *
 * A "band" is rows [row0, row0+rows) of every layer of every frame, resident
 * as planar WORD [n_frames][layers][buf_rows][width] (frame stride and layer
 * stride given in elements, so whole frames resident in HBM are the case
 * buf_row0 = 0, buf_rows = height). */
typedef struct ssc_band {
	const WORD *frames;		/* device: frame f, layer c, storage row r, column x at
					   frames[f*frame_stride + c*layer_stride + (r-buf_row0)*width + x] */
	size_t frame_stride;		/* elements */
	size_t layer_stride;		/* elements */
	int n_frames, layers, width;
	int height;			/* full image height */
	int buf_row0, buf_rows;		/* storage rows present in the buffer */
	int out_row0, out_rows;		/* result storage rows this call produces */
	int row_order;			/* SSC_ROWS_* */
	const int *shiftx, *shifty;	/* device int[n_frames] or NULL */
} ssc_band;

/* Counters written by ssc_dev_stack_mean (device memory, uint64_t[8]):
 * [0..5] irej[3][2], [6] guard_fired, [7] cap_hit.  The caller zeroes them. */
#define SSC_COUNTERS 8

/* K2: stack_median hot loop (stacking.c:746-769).  coeff = device double[3][n_frames]
 * (offset, mul, scale) or NULL for identity, as compute_normalization / ssc_dev_norm_coeffs
 * produce them: scale is 1.0 in the ADDITIVE and MULTIPLICATIVE modes (stacking.c:130-134).  out: device WORD [layers][out_pitch_rows][width]
 * addressed by result storage row - out_buf_row0. */
int ssc_dev_stack_median(const ssc_band *band, int normalize, const double *coeff,
		WORD *out, size_t out_layer_stride, int out_buf_row0, void *stream);

/* K3: stack_mean_with_rejection hot loop (stacking.c:1620-1802). */
int ssc_dev_stack_mean(const ssc_band *band, int normalize, const double *coeff,
		int type_of_rejection, double sig_low, double sig_high,
		WORD *out, size_t out_layer_stride, int out_buf_row0,
		uint64_t *counters, void *stream);

/* K1: stack_summing accumulation (stacking.c:298-319): acc (device uint32_t
 * [layers][out_rows][width]) = per-pixel sums, *maxim (device uint64_t, caller
 * zeroes) = atomicMax of the sums.  K1b: scale pass (:328-341) with the global max. */
int ssc_dev_sum_accumulate(const ssc_band *band, uint32_t *acc, unsigned long long *maxim, void *stream);
int ssc_dev_sum_finish(const uint32_t *acc, size_t count, const unsigned long long *maxim,
		WORD *out, void *stream);

/* K1c: stack_addmax (take_max != 0) / stack_addmin in one pass (stacking.c:918-936, 1074-1092):
 * out (device WORD [layers][out_rows][width]) = brightest / darkest sample per pixel;
 * *extreme (device unsigned int; the caller sets 0 for max, 0xFFFF for min before the first
 * band of an image) folds the max / min of the result (gfit.hi). */
int ssc_dev_stack_extreme(const ssc_band *band, int take_max, WORD *out, unsigned int *extreme, void *stream);

/* K4: exact 65536-bin histograms of layer 0 of every frame of the band
 * (replaces the sort in statistics.c:279-285).  hist: device uint32_t
 * [n_frames][SSC_HIST_STRIDE], caller zeroes; bins [0,65536) are the value
 * counts, word 65536 counts samples > 255 in layers 1.. (only gathered for
 * frames whose layer 0 stays <= 255 in this band; decides hist_size 256 vs
 * 65536, gui/histogram.c:116).  Histograms of several bands / GPUs add up
 * (integer SUM all-reduce) before ssc_dev_ikss. */
#define SSC_HIST_STRIDE 65540
int ssc_dev_histogram(const ssc_band *band, uint32_t *hist, void *stream);

/* K5: IKSS location/scale of every frame from its (combined) histogram
 * (statistics.c:152-187,277-290).  stats: device double[4][n_frames] =
 * location, scale, normValue, ngoodpix.  scratch: device uint32_t
 * [n_frames][65536]. */
int ssc_dev_ikss(const uint32_t *hist, int n_frames, double *stats, uint32_t *scratch, void *stream);

/* K6: _compute_normalization_for_image (stacking.c:79-123) for all frames:
 * coeff device double[3][n_frames] = offset, mul, scale.  status (device int,
 * may be NULL) is set to 1 when a frame has no non-zero pixel. */
int ssc_dev_norm_coeffs(const double *stats, int n_frames, int ref_image, int mode,
		double *coeff, int *status, void *stream);

/* ------------------------------------------------------------------------- */
/* Several GPUs: the exchanges of the path as kernels over NVLink peer memory */
/* ------------------------------------------------------------------------- */
/* The image is split in row bands, one per GPU ("rank"); what the reference keeps in shared
 * variables of its OpenMP threads becomes three small exchanges: the per-frame
 * histograms add up before the IKSS (stacking.c:166-187 computes them per frame on whole images),
 * stack_summing's maximum (stacking.c:311-313) and the rejection counters (:1796-1802).  Every rank
 * owns a block of device memory that all others map; the kernels add / store into it directly, no
 * host in between.  One process per GPU: ssc_peer_export / ssc_peer_import (cudaIpc handles, passed
 * around by the host's own means); one process driving several GPUs: ssc_peer_attach_local.  All
 * ranks make the same calls in the same order; `stream` as in the device layer. */
typedef struct ssc_peer_group ssc_peer_group;
#define SSC_PEER_MAX 64
#define SSC_PEER_HANDLE_BYTES 64
#define SSC_PEER_TIMEOUT (-110)	/* a rank waited ~10 s for a peer that never arrived (ssc_peer_status) */
/* on the current device; n_frames_max bounds n_frames of the later calls */
int ssc_peer_create(int world, int rank, int n_frames_max, ssc_peer_group **out);
int ssc_peer_export(ssc_peer_group *g, void *handle64);
int ssc_peer_import(ssc_peer_group *g, int peer_rank, const void *handle64);
int ssc_peer_attach_local(ssc_peer_group *g, int peer_rank, ssc_peer_group *other);
int ssc_peer_connect(ssc_peer_group *g);		/* after every peer was imported / attached */
int ssc_peer_destroy(ssc_peer_group *g);
int ssc_peer_world(const ssc_peer_group *g);
int ssc_peer_rank(const ssc_peer_group *g);
int ssc_peer_status(ssc_peer_group *g, void *stream);	/* 0 or SSC_PEER_TIMEOUT; synchronises the stream */
const char *ssc_peer_last_error(void);
/* K4 fused with the reduce-scatter of the histograms: frame f is analysed by rank f % world and this
 * band's counts are added straight into that rank's histogram.  last_piece = 0 when more row pieces
 * of this rank's band follow in the same exchange. */
int ssc_dev_histogram_scatter(const ssc_band *band, ssc_peer_group *g, int last_piece, void *stream);
/* K5 on this rank's frames once all ranks' counts are in, then the all-gather of the 4 statistics per
 * frame.  scratch: device uint32_t [ceil(n_frames / world)][65536]. */
int ssc_dev_ikss_gather(ssc_peer_group *g, int n_frames, uint32_t *scratch, void *stream);
/* K6 once every owner's statistics are here: identical coefficients on every rank.  stats_out (device
 * double[4][n_frames]) may be NULL. */
int ssc_dev_norm_coeffs_gathered(ssc_peer_group *g, int n_frames, int ref_image, int mode, double *coeff,
		int *status, double *stats_out, void *stream);
/* total (device uint64_t[SSC_COUNTERS]) = sum over the ranks of local (stacking.c:1796-1802) */
int ssc_dev_counters_combine(ssc_peer_group *g, const uint64_t *local, uint64_t *total, void *stream);
/* *maxim (device uint64_t) = max over the ranks, between K1 and K1b (stacking.c:311-313, 328-330) */
int ssc_dev_max_combine(ssc_peer_group *g, unsigned long long *maxim, void *stream);

/* ------------------------------------------------------------------------- */
/* File formats either side of the path                 */
/* ------------------------------------------------------------------------- */
/* What readfits learns from a FITS header (image_format_fits.c:40-130), from the first bytes of the file
 * (len = bytes available; the whole file or at least the header). */
typedef struct ssc_fits_info {
	int bitpix, naxis;
	long naxes[3];			/* rx, ry, layers (1 or 3) */
	double bzero, bscale, exposure;	/* EXPTIME / EXPOSURE */
	WORD hi, lo;			/* MIPS-HI / MIPS-LO */
	size_t data_offset, data_bytes;	/* the data unit */
	int frame_format;		/* SSC_FRAMES_FITS16 or SSC_FRAMES_RAW8 */
} ssc_fits_info;
int ssc_fits_parse_header(const void *bytes, size_t len, ssc_fits_info *info);
/* ser_read_header (ser.c:291-350) from the 178 header bytes; file_size fixes a missing frame count (:336-345) */
typedef struct ssc_ser_info {
	int color_id, big_endian, width, height, bit_depth, frame_count;
	int bytes_per_sample, planes;
	size_t frame_bytes, data_offset;	/* frame i starts at data_offset + i * frame_bytes */
	int frame_format;		/* SSC_FRAMES_* of its frames; row order SSC_ROWS_SER */
} ssc_ser_info;
int ssc_ser_parse_header(const void *bytes, size_t file_size, ssc_ser_info *info);
/* savefits (image_format_fits.c:652-739) of a stacking result: BITPIX 16, BZERO 32768, samples encoded on the
 * device.  data: host, planar [layers][ry][rx] (ssc_result.data). */
int ssc_savefits(const char *path, const WORD *data, int rx, int ry, int layers, WORD hi, WORD lo, double exposure,
		int overwrite);
/* statistics(fit, layer, NULL, STATS_BASIC, STATS_ZERO_NULLCHECK) (src/algos/statistics.c:207-326) for every layer
 * of an image, as noise() (src/core/siril.c:1740-1771) and the stacking summary run it on the result: the fields of
 * imstats (src/core/siril.h:600-605) that STATS_BASIC fills.  data: host, planar [layers][ry][rx]. */
typedef struct ssc_imstats {
	double total, ngoodpix, mean, median, sigma, bgnoise, min, max, normValue;
} ssc_imstats;
int ssc_image_statistics(const WORD *data, int rx, int ry, int layers, int nullcheck, ssc_imstats *stats);
const char *ssc_io_last_error(void);
/* the decoders and preprocess on samples that are already on the device (what the operators run behind their uploads) */
int ssc_dev_decode_fits16(WORD *samples, size_t count, void *stream);		/* in place */
int ssc_dev_decode_ser(const void *raw, WORD *planes, size_t layer_stride, size_t pixels, int bytes_per_sample,
		int big_endian, int color_id, void *stream);
int ssc_dev_preprocess(WORD *samples, size_t count, const WORD *offset, const WORD *dark, const WORD *flat,
		float level, void *stream);

/* Counter-based synthetic frames for benchmarks and large-size
 * property tests: fills device WORD [n_frames][layers][rows][width] with storage
 * rows [row0, row0+rows) of frames first_frame.. of a full_height-row image; every
 * sample depends only on (seed, kind, frame, layer, y, x), so any band on any GPU
 * regenerates the same pixels.  kind 0 = deep-sky field, 1 = planetary disc. */
int ssc_dev_synth_frames(WORD *frames, int n_frames, int layers, int rows, int width,
		int first_frame, int row0, int full_height, int kind, uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SIRIL_STACK_CUDA_H */
