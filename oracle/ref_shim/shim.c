/* oracle/ref_shim/shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the
 * product library; only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load oracle/_ref/libsiril_ref.so).
 *
 * Everything the reference's UNMODIFIED src/stacking/stacking.c and
 * src/algos/statistics.c need from the rest of Siril, reduced to an in-memory
 * frame feed:
 *   - the globals com, gfit, wfit (src/main.c:56-62);
 *   - the sequence feed functions of src/io/sequence.c (seq_read_frame :519,
 *     seq_open_image :635, seq_close_image :674, seq_opened_read_region :690,
 *     seq_get_imstats :1107, sequence_is_rgb :1089) backed by a planar
 *     u16 [N][C][H][W] array instead of FITS/SER files;
 *   - the three src/core/utils.c primitives (round_to_WORD :68, quicksort_d
 *     :488, quicksort_s :512) restated (utils.c itself drags in GTK CSS code);
 *   - computeHisto (src/gui/histogram.c:111) + get_normalized_value
 *     (src/core/utils.c:454) + image_find_minmax (src/io/single_image.c:220);
 *   - copyfits/clearfits (src/io/image_format_fits.c:958) for the flags used;
 *   - cfitsio's fits_img_stats_ushort reduced to what the path reads (ngoodpix);
 *   - log / progress / cancel hooks as no-ops (log lines are parsed to recover
 *     the rejection counters, which the reference keeps in a local variable,
 *     stacking.c:1193,1811-1817).
 *
 * The fake sequence is SEQ_SER-typed because that branch of the stackers takes
 * the frame geometry from seq->ser_file and never touches cfitsio
 * (stacking.c:484-499,1311-1326).  The ROW ORDER seen by the stackers is
 * selectable (oref_config.row_order):
 *   0 = FITS convention: read_opened_fits_partial hands rows top-down out of a
 *       bottom-up file (image_format_fits.c:599-632), the stacker writes block
 *       row y to ry-y-1 (stacking.c:731,1600): storage row r -> storage row r.
 *   1 = SER convention: ser_read_opened_partial hands rows in file order
 *       (ser.c:797-799) and ser_read_frame flips whole frames (ser.c:767):
 *       the result is vertically flipped with respect to the input storage.
 */
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <sys/time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAIN	/* we define the globals that siril.h would otherwise declare extern */
#include "core/siril.h"
#include "core/proto.h"
#include "io/sequence.h"
#include "io/ser.h"
#include "gui/callbacks.h"
#include "gui/histogram.h"
#include "registration/registration.h"
#include "stacking/stacking.h"
#include "oref_api.h"

/* ---- globals of src/main.c ---- */
cominfo com;
fits gfit;
fits wfit[5];

/* ---- the in-memory sequence ---- */
static struct {
	const WORD *frames;	/* [N][C][H][W] */
	int n, w, h, c;
	int row_order;
	int reglayer;
	uint64_t irej[3][2];
	int verbose;
} feed;

static const WORD *feed_plane(int index, int layer) {
	size_t plane = (size_t) feed.w * feed.h;
	return feed.frames + ((size_t) index * feed.c + layer) * plane;
}

int get_registration_layer() {
	return feed.reglayer;
}

gboolean get_thread_run() {
	return TRUE;
}

void set_progress_bar_data(const char *text, double percent) {
	(void) text;
	(void) percent;
}

/* returns a heap string the callers may poke at (stacking.c:139-140) */
static char *oracle_log(const char *format, va_list ap) {
	static __thread char msg[1024];
	va_list ap2;
	va_copy(ap2, ap);
	if (!strncmp(format, "Pixel rejection in channel", 26)) {
		/* stacking.c:1811-1817: args are (long channel, double low%, double high%)
		 * with low% = irej/(W*H*N)*100 -- invert it, the quotient is far from
		 * half-integers so rint() recovers the exact counter */
		long channel = va_arg(ap2, long);
		double lo = va_arg(ap2, double), hi = va_arg(ap2, double);
		double nb_tot = (double) feed.w * feed.h * feed.n;
		if (channel >= 0 && channel < 3) {
			feed.irej[channel][0] = (uint64_t) llrint(lo / 100.0 * nb_tot);
			feed.irej[channel][1] = (uint64_t) llrint(hi / 100.0 * nb_tot);
		}
	}
	va_end(ap2);
	vsnprintf(msg, sizeof(msg), format, ap);
	if (feed.verbose)
		fputs(msg, stderr);
	return msg;
}

char *siril_log_message(const char *format, ...) {
	va_list ap;
	char *r;
	va_start(ap, format);
	r = oracle_log(format, ap);
	va_end(ap);
	return r;
}

char *siril_log_color_message(const char *format, const char *color, ...) {
	va_list ap;
	char *r;
	(void) color;
	va_start(ap, color);
	r = oracle_log(format, ap);
	va_end(ap);
	return r;
}

int get_available_memory_in_MB() {
	FILE *f = fopen("/proc/meminfo", "r");
	char line[256];
	long kb = 0;
	if (!f) return 2048;
	while (fgets(line, sizeof(line), f))
		if (sscanf(line, "MemAvailable: %ld kB", &kb) == 1)
			break;
	fclose(f);
	return kb > 0 ? (int) (kb / 1024) : 2048;
}

char *str_append(char **data, const char *newdata) {
	size_t len = strlen(*data) + strlen(newdata) + 1;
	char *p = realloc(*data, len);
	if (!p) return NULL;
	strcat(p, newdata);
	*data = p;
	return p;
}

/* ---- src/core/utils.c primitives, restated ---- */
WORD round_to_WORD(double x) {
	if (x <= 0.0) return 0;
	if (x > 65535.0) return 65535;
	return (WORD) (x + 0.5);
}

/* Same algorithm class as the reference (recursive Hoare partition around the
 * middle element) so that the CPU baseline pays the reference's sort cost. */
#define ORACLE_QSORT(NAME, T) \
void NAME(T *a, int n) { \
	while (n > 1) { \
		T pivot = a[n / 2]; \
		int lo = 0, hi = n - 1; \
		while (lo <= hi) { \
			while (a[lo] < pivot) lo++; \
			while (a[hi] > pivot) hi--; \
			if (lo <= hi) { T t = a[lo]; a[lo] = a[hi]; a[hi] = t; lo++; hi--; } \
		} \
		NAME(a, hi + 1); \
		a += lo; n -= lo; \
	} \
}
ORACLE_QSORT(quicksort_s, WORD)
ORACLE_QSORT(quicksort_d, double)

/* ---- fits helpers ---- */
void clearfits(fits *fit) {
	if (!fit) return;
	free(fit->data);
	free(fit->header);
	memset(fit, 0, sizeof(fits));
}

int copyfits(fits *from, fits *to, unsigned char oper, int layer) {
	unsigned int nbdata = from->rx * from->ry;
	int depth = (oper & CP_EXPAND) ? 3 : (int) from->naxes[2];
	(void) layer;
	if (oper & CP_ALLOC) {
		WORD *p = realloc(to->data, (size_t) nbdata * depth * sizeof(WORD));
		if (!p) { free(to->data); to->data = NULL; return -1; }
		to->data = p;
		to->pdata[0] = p;
		to->pdata[1] = depth == 3 ? p + nbdata : p;
		to->pdata[2] = depth == 3 ? p + 2 * (size_t) nbdata : p;
	}
	if (oper & CP_INIT)
		memset(to->data, 0, (size_t) nbdata * depth * sizeof(WORD));
	if (oper & CP_COPYA)
		memcpy(to->data, from->data, (size_t) nbdata * depth * sizeof(WORD));
	if (oper & CP_FORMAT) {
		to->rx = from->rx; to->ry = from->ry;
		to->lo = from->lo; to->hi = from->hi;
		to->bitpix = from->bitpix;
		to->naxis = depth == 3 ? 3 : from->naxis;
		to->naxes[0] = from->naxes[0];
		to->naxes[1] = from->naxes[1];
		to->naxes[2] = depth;
	}
	return 0;
}

/* ---- sequence feed ---- */
char *seq_get_image_filename(sequence *seq, int index, char *name_buf) {
	(void) seq;
	sprintf(name_buf, "mem_%05d", index);
	return name_buf;
}

char *fit_sequence_get_image_filename(sequence *seq, int index, char *name_buffer, gboolean add_fits_ext) {
	(void) add_fits_ext;
	return seq_get_image_filename(seq, index, name_buffer);
}

int seq_read_frame(sequence *seq, int index, fits *dest) {
	size_t plane = (size_t) feed.w * feed.h;
	int layer, y;
	(void) seq;
	if (index < 0 || index >= feed.n) return 1;
	free(dest->data);
	memset(dest, 0, sizeof(fits));
	dest->data = malloc(plane * feed.c * sizeof(WORD));
	if (!dest->data) return 1;
	dest->rx = feed.w; dest->ry = feed.h;
	dest->naxes[0] = feed.w; dest->naxes[1] = feed.h; dest->naxes[2] = feed.c;
	dest->naxis = feed.c == 3 ? 3 : 2;
	dest->bitpix = USHORT_IMG;
	dest->exposure = 1.0;
	for (layer = 0; layer < 3; layer++)
		dest->pdata[layer] = dest->data + (feed.c == 3 ? layer * plane : 0);
	for (layer = 0; layer < feed.c; layer++) {
		const WORD *src = feed_plane(index, layer);
		if (feed.row_order == 0)
			memcpy(dest->pdata[layer], src, plane * sizeof(WORD));
		else	/* ser_read_frame ends with fits_flip_top_to_bottom (ser.c:767) */
			for (y = 0; y < feed.h; y++)
				memcpy(dest->pdata[layer] + (size_t) y * feed.w,
						src + (size_t) (feed.h - 1 - y) * feed.w, feed.w * sizeof(WORD));
	}
	return 0;
}

int seq_open_image(sequence *seq, int index) {
	(void) seq; (void) index;
	return 0;
}

void seq_close_image(sequence *seq, int index) {
	(void) seq; (void) index;
}

int seq_opened_read_region(sequence *seq, int layer, int index, WORD *buffer, const rectangle *area) {
	const WORD *src;
	int j;
	(void) seq;
	if (index < 0 || index >= feed.n || layer < 0 || layer >= feed.c) return -1;
	if (area->x < 0 || area->y < 0 || area->x + area->w > feed.w || area->y + area->h > feed.h)
		return -1;
	src = feed_plane(index, layer);
	for (j = 0; j < area->h; j++) {
		int srow = feed.row_order == 0 ? feed.h - 1 - (area->y + j) : area->y + j;
		memcpy(buffer + (size_t) j * area->w, src + (size_t) srow * feed.w + area->x,
				area->w * sizeof(WORD));
	}
	return 0;
}

gboolean sequence_is_rgb(sequence *seq) {
	(void) seq;
	return feed.c == 3;
}

imstats *seq_get_imstats(sequence *seq, int index, fits *the_image, int option) {
	if (!seq->imgparam[index].stats && the_image) {
		seq->imgparam[index].stats = statistics(the_image, 0, NULL, option, STATS_ZERO_NULLCHECK);
		if (!seq->imgparam[index].stats)
			return NULL;
		seq->needs_saving = TRUE;
	}
	return seq->imgparam[index].stats;
}

/* ---- histogram front end of statistics() ---- */
gsl_histogram *computeHisto(fits *fit, int layer) {
	size_t i, ndata = (size_t) fit->rx * fit->ry, size;
	int l;
	gsl_histogram *histo;
	if (fit->maxi == 0) {	/* image_find_minmax(fit, 0) */
		WORD mx = 0, mn = 65535;
		for (l = 0; l < fit->naxes[2]; l++) {
			const WORD *b = fit->pdata[l];
			WORD lmx = 0, lmn = 65535;
			for (i = 0; i < ndata; i++) {
				if (b[i] > lmx) lmx = b[i];
				if (b[i] < lmn) lmn = b[i];
			}
			fit->max[l] = lmx; fit->min[l] = lmn;
			if (lmx > mx) mx = lmx;
			if (lmn < mn) mn = lmn;
		}
		fit->maxi = mx; fit->mini = mn;
	}
	size = fit->maxi <= 255 ? 255 : 65535;	/* get_normalized_value */
	histo = gsl_histogram_alloc(size + 1);
	gsl_histogram_set_ranges_uniform(histo, 0, size);
	for (i = 0; i < ndata; i++)
		gsl_histogram_increment(histo, (double) fit->pdata[layer][i]);
	return histo;
}

gsl_histogram *computeHisto_Selection(fits *fit, int layer, rectangle *selection) {
	(void) selection;
	return computeHisto(fit, layer);
}

/* ---- cfitsio ---- */
int fits_img_stats_ushort(unsigned short *array, long nx, long ny, int nullcheck,
		unsigned short nullvalue, long *ngoodpix, unsigned short *minvalue,
		unsigned short *maxvalue, double *mean, double *sigma, double *noise1,
		double *noise2, double *noise3, double *noise5, int *status) {
	long i, n = nx * ny, good = 0;
	double s = 0, s2 = 0;
	(void) minvalue; (void) maxvalue; (void) noise2; (void) noise3; (void) noise5;
	for (i = 0; i < n; i++) {
		if (nullcheck && array[i] == nullvalue) continue;
		good++;
		s += array[i];
		s2 += (double) array[i] * array[i];
	}
	if (ngoodpix) *ngoodpix = good;
	if (mean) *mean = good ? s / good : 0;
	if (sigma) *sigma = good > 1 ? sqrt(fmax(0, (s2 - s * s / good) / (good - 1))) : 0;
	if (noise1) *noise1 = 0;	/* not read by the stacking path */
	return *status;
}

static void oracle_unreachable(const char *what) {
	fprintf(stderr, "oracle shim: %s must not be reached with a SEQ_SER feed\n", what);
	abort();
}
int fits_get_img_param(fitsfile *f, int maxaxis, int *bitpix, int *naxis, long *naxes, int *status) {
	(void) f; (void) maxaxis; (void) bitpix; (void) naxis; (void) naxes; (void) status;
	oracle_unreachable("fits_get_img_param"); return 1;
}
int fits_read_key(fitsfile *f, int datatype, const char *keyname, void *value, char *comm, int *status) {
	(void) f; (void) datatype; (void) keyname; (void) value; (void) comm; (void) status;
	oracle_unreachable("fits_read_key"); return 1;
}
int fits_is_reentrant(void) { return 1; }
void fits_report_error(FILE *stream, int status) { (void) stream; (void) status; }

/* ---- public entry points (oref_api.h) ---- */
static int oref_filter_all(sequence *seq, int nb_img, double any) {
	(void) seq; (void) nb_img; (void) any;
	return 1;
}

static double now_s(void) {
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static int last_hi = -1;
/* gfit.hi as the last successful operator left it (sum, addmax, addmin set it) */
int oref_last_hi(void) {
	return last_hi;
}

int oref_num_procs(void) {
#ifdef _OPENMP
	return omp_get_num_procs();
#else
	return 1;
#endif
}

int oref_run(const oref_config *cfg, unsigned short *out, uint64_t irej[6],
		double *coeff /* [3][N] offset,mul,scale or NULL */, double timing[2]) {
	sequence *seq = &com.seq;
	struct ser_struct ser;
	struct stacking_args args;
	int i, retval, nthreads;
	int *indices;
	regdata *reg = NULL;
	double t0, t1;
	size_t npix = (size_t) cfg->width * cfg->height * cfg->channels;

	memset(&feed, 0, sizeof(feed));
	feed.frames = cfg->frames;
	feed.n = cfg->n_frames; feed.w = cfg->width; feed.h = cfg->height; feed.c = cfg->channels;
	feed.row_order = cfg->row_order;
	feed.reglayer = (cfg->shiftx || cfg->shifty) ? 0 : -1;
	feed.verbose = cfg->verbose;

	nthreads = cfg->max_thread > 0 ? cfg->max_thread : oref_num_procs();	/* main.c:373-379 */
	memset(&com, 0, sizeof(com));
	com.max_thread = nthreads;
	com.run_thread = TRUE;
	com.stack.memory_percent = 0.9;

	memset(&ser, 0, sizeof(ser));
	ser.color_id = cfg->channels == 3 ? SER_RGB : SER_MONO;
	ser.image_width = cfg->width;
	ser.image_height = cfg->height;
	ser.bit_pixel_depth = 16;
	ser.frame_count = cfg->n_frames;
	ser.byte_pixel_depth = SER_PIXEL_DEPTH_16;
	ser.number_of_planes = cfg->channels;

	seq->seqname = "oracle";
	seq->number = cfg->n_frames;
	seq->selnum = cfg->n_frames;
	seq->nb_layers = cfg->channels;
	seq->rx = cfg->width; seq->ry = cfg->height;
	seq->reference_image = cfg->reference_image;
	seq->type = SEQ_SER;
	seq->ser_file = &ser;
	seq->imgparam = calloc(cfg->n_frames, sizeof(imgdata));
	seq->regparam = calloc(3, sizeof(regdata *));
	if (feed.reglayer >= 0) {
		reg = calloc(cfg->n_frames, sizeof(regdata));
		for (i = 0; i < cfg->n_frames; i++) {
			reg[i].shiftx = cfg->shiftx ? cfg->shiftx[i] : 0;
			reg[i].shifty = cfg->shifty ? cfg->shifty[i] : 0;
		}
		seq->regparam[0] = reg;
	}
	indices = malloc(cfg->n_frames * sizeof(int));
	for (i = 0; i < cfg->n_frames; i++) {
		indices[i] = i;
		seq->imgparam[i].incl = TRUE;
		seq->imgparam[i].filenum = i;
	}

	memset(&args, 0, sizeof(args));
	args.seq = seq;
	args.filtering_criterion = oref_filter_all;
	args.filtering_parameter = 0.0;
	args.nb_images_to_stack = cfg->n_frames;
	args.image_indices = indices;
	args.sig[0] = cfg->sig_low;
	args.sig[1] = cfg->sig_high;
	args.type_of_rejection = (rejection) cfg->rejection;
	args.normalize = (normalization) cfg->normalize;
	args.force_norm = TRUE;
	if (cfg->max_number_of_rows > 0) {
		args.max_number_of_rows = cfg->max_number_of_rows;
	} else {
		/* stacking.c:1903-1915 */
		int max_memory = (int) (com.stack.memory_percent * (double) get_available_memory_in_MB());
		uint64_t rows = (uint64_t) max_memory * 1048576L /
			((uint64_t) seq->rx * args.nb_images_to_stack * sizeof(WORD) * com.max_thread);
		if (rows > seq->ry) args.max_number_of_rows = seq->ry;
		else if (rows * 2 > seq->ry) args.max_number_of_rows = seq->ry / 2;
		else args.max_number_of_rows = (int) rows;
	}

	memset(&gfit, 0, sizeof(gfit));
	memset(wfit, 0, sizeof(wfit));
	timing[0] = timing[1] = 0.0;

	if (cfg->method == OREF_NORMALIZATION_ONLY || (coeff && cfg->normalize != NO_NORM)) {
		/* separately timed compute_normalization (stacking.c:125); the stackers
		 * then hit the imgparam[].stats cache instead of recomputing */
		norm_coeff nc;
		nc.offset = coeff ? coeff : malloc(cfg->n_frames * sizeof(double));
		nc.mul = coeff ? coeff + cfg->n_frames : malloc(cfg->n_frames * sizeof(double));
		nc.scale = coeff ? coeff + 2 * cfg->n_frames : malloc(cfg->n_frames * sizeof(double));
		t0 = now_s();
		retval = compute_normalization(&args, &nc, args.normalize);
		timing[1] = now_s() - t0;
		if (!coeff) { free(nc.offset); free(nc.mul); free(nc.scale); }
		args.force_norm = FALSE;
		if (retval || cfg->method == OREF_NORMALIZATION_ONLY)
			goto done;
	}

	switch (cfg->method) {
	case OREF_SUM: args.method = stack_summing; break;
	case OREF_MEAN: args.method = stack_mean_with_rejection; break;
	case OREF_MEDIAN: args.method = stack_median; break;
	case OREF_ADDMAX: args.method = stack_addmax; break;
	case OREF_ADDMIN: args.method = stack_addmin; break;
	default: retval = -100; goto done;
	}
	t0 = now_s();	/* same bracket as stacking.c:1918 -> :2102-2103 */
	retval = args.method(&args);
	t1 = now_s();
	timing[0] = t1 - t0;

	if (!retval && gfit.data && out)
		memcpy(out, gfit.data, npix * sizeof(WORD));
	last_hi = retval ? -1 : (int) gfit.hi;
	if (irej)
		for (i = 0; i < 6; i++)
			irej[i] = feed.irej[i / 2][i % 2];
done:
	free(gfit.data); gfit.data = NULL;
	for (i = 0; i < 5; i++) { free(wfit[i].data); wfit[i].data = NULL; }
	for (i = 0; i < cfg->n_frames; i++)
		free(seq->imgparam[i].stats);
	free(seq->imgparam); seq->imgparam = NULL;
	free(seq->regparam); seq->regparam = NULL;
	free(reg);
	free(indices);
	return retval;
}

/* location/scale of one planar frame exactly as normalization sees them
 * (sequence.c:1110 -> statistics(fit, 0, NULL, STATS_EXTRA, nullcheck)) */
int oref_frame_stats(const unsigned short *frame, int width, int height, int channels,
		double *location, double *scale, double *norm_value, long *ngoodpix) {
	fits fit;
	imstats *st;
	size_t plane = (size_t) width * height;
	int l;
	memset(&fit, 0, sizeof(fit));
	fit.rx = width; fit.ry = height;
	fit.naxes[0] = width; fit.naxes[1] = height; fit.naxes[2] = channels;
	fit.naxis = channels == 3 ? 3 : 2;
	fit.bitpix = USHORT_IMG;
	fit.data = (WORD *) frame;
	for (l = 0; l < 3; l++)
		fit.pdata[l] = fit.data + (channels == 3 ? l * plane : 0);
	st = statistics(&fit, 0, NULL, STATS_EXTRA, STATS_ZERO_NULLCHECK);
	if (!st) return -1;
	*location = st->location;
	*scale = st->scale;
	*norm_value = st->normValue;
	*ngoodpix = st->ngoodpix;
	free(st);
	return 0;
}
