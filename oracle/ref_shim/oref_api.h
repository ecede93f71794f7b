/* oracle/ref_shim/oref_api.h -- TEST INFRASTRUCTURE ONLY.
 * ctypes-friendly entry points of oracle/_ref/libsiril_ref.so: the reference's
 * own stacking.c / statistics.c (compiled unmodified from /root/reference) run
 * on an in-memory planar u16 [N][C][H][W] sequence. */
#ifndef OREF_API_H
#define OREF_API_H
#include <stdint.h>

enum { OREF_SUM = 0, OREF_MEAN = 1, OREF_MEDIAN = 2, OREF_NORMALIZATION_ONLY = 3, OREF_ADDMAX = 4, OREF_ADDMIN = 5 };

typedef struct {
	const unsigned short *frames;	/* [N][C][H][W] in storage order */
	int n_frames, width, height, channels;
	int method;		/* OREF_* (index order of stacking_methods[], stacking.c:54-56) */
	int rejection;		/* reference enum, stacking.h:14-21 */
	int normalize;		/* reference enum, stacking.h:24-30 */
	double sig_low, sig_high;
	int reference_image;	/* seq->reference_image, -1 = unset */
	int row_order;		/* 0 FITS convention, 1 SER convention (see shim.c) */
	const int *shiftx, *shifty;	/* regparam[reglayer][i], NULL = no registration */
	int max_thread;		/* com.max_thread; <=0: omp_get_num_procs() */
	int max_number_of_rows;	/* <=0: the reference's formula, stacking.c:1903-1915 */
	int verbose;
} oref_config;

int oref_num_procs(void);
int oref_last_hi(void);	/* gfit.hi after the last successful oref_run, -1 otherwise */
/* out: [C][H][W]; irej: [3][2] low/high; coeff: [3][N] offset,mul,scale (may be NULL);
 * timing[0] = seconds in args->method(args), timing[1] = seconds in a separate
 * compute_normalization call (0 if not run separately). Returns the operator's retval. */
int oref_run(const oref_config *cfg, unsigned short *out, uint64_t irej[6],
		double *coeff, double timing[2]);
int oref_frame_stats(const unsigned short *frame, int width, int height, int channels,
		double *location, double *scale, double *norm_value, long *ngoodpix);
#endif
