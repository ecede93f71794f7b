/* oracle/siril_oracle.h -- CPU restatement of Siril 0.9.7's per-pixel stacking
 * hot path.  TEST INFRASTRUCTURE ONLY: it exists to check the CUDA library
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) and is never
 * linked into, loaded by or shipped with libsiril_stack_cuda.
 *
 * PINNING: the reference ships no tests, fixtures or golden vectors for this
 * path.  This restatement is pinned instead against the
 * reference's own stacking.c / statistics.c compiled unmodified into
 * oracle/_ref/libsiril_ref.so (oracle/Makefile) -- tests/test_oracle_pinning.py
 * compares the two on seeded inputs, and tests/golden/ holds outputs generated
 * from oracle/_ref by tests/golden/make_golden.py.  GSL itself is not under
 * /root/reference (un-vendored, unpinned version): its three routines are
 * restated from the published algorithms in both oracles, so GSL fidelity is
 * "parity unpinned"; everything on the Siril side is pinned.
 *
 * All images are planar uint16 [N][C][H][W]; results are planar [C][H][W].
 */
#ifndef SIRIL_ORACLE_H
#define SIRIL_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference enums, src/stacking/stacking.h:14-30 */
enum { SO_NO_REJEC = 0, SO_PERCENTILE, SO_SIGMA, SO_SIGMEDIAN, SO_WINSORIZED, SO_LINEARFIT };
enum { SO_NO_NORM = 0, SO_ADDITIVE, SO_MULTIPLICATIVE, SO_ADDITIVE_SCALING, SO_MULTIPLICATIVE_SCALING };

/* how sigma (gsl_stats_ushort_sd) is evaluated */
enum {
	SO_SD_GSL_LONG_DOUBLE = 0,	/* GSL's running mean/variance in x87 long double (the reference) */
	SO_SD_EXACT_INTEGER = 1,	/* sqrt(((n*Q-S*S)/(n*n)) * (n/(n-1))) from exact integer sums (the GPU) */
	/* flag, OR-ed into sd_mode: how the straight line of LINEARFIT (gsl_fit_linear + the residual sum)
	 * is evaluated.  Clear: GSL's running means, literally (the reference).  Set: from the exact integer
	 * sums of y and i*y with one rounding for slope and intercept, and the mean absolute residual from
	 * an exact fixed-point sum (order independent) -- the GPU's form, see fit_exact() */
	SO_FIT_EXACT_SUMS = 2
};

typedef struct {
	const uint16_t *frames;		/* [N][C][H][W] */
	int n_frames, width, height, channels;
	int normalize;			/* SO_*NORM* */
	int rejection;			/* SO_* rejection */
	double sig_low, sig_high;	/* args->sig[0], args->sig[1] */
	const double *coeff;		/* [3][N]: offset, mul, scale; NULL = identity */
	const int *shiftx, *shifty;	/* per frame; both NULL = no registration data */
	int row_order;			/* 0 FITS convention (identity), 1 SER convention (vertical flip) */
	int sd_mode;			/* SO_SD_* */
	int threads;			/* <=0: all */
} so_stack_args;

uint16_t so_round_to_word(double x);

/* stack_summing, stacking.c:196-355.  maxim = the global maximum sum. */
int so_stack_sum(const so_stack_args *a, uint16_t *out, uint64_t *maxim);
/* stack_addmax (take_max != 0, stacking.c:824-972) / stack_addmin (stacking.c:979-1128).
 * extreme = the value the reference stores in gfit.hi. */
int so_stack_extreme(const so_stack_args *a, int take_max, uint16_t *out, uint16_t *extreme);
/* stack_median, stacking.c:362-816 (hot loop :746-769) */
int so_stack_median(const so_stack_args *a, uint16_t *out);
/* stack_mean_with_rejection, stacking.c:1189-1858 (hot loop :1620-1802).
 * irej[3][2]; guard_fired = pixels where the "N - r <= 4" break was taken;
 * cap_hit = pixels whose do-while ran into SO_ITER_CAP. */
int so_stack_mean(const so_stack_args *a, uint16_t *out, uint64_t irej[6],
		uint64_t *guard_fired, uint64_t *cap_hit);
/* same, plus a per-pixel map [C][H][W] of the guard events (1 = history dependent in the reference) */
int so_stack_mean_map(const so_stack_args *a, uint16_t *out, uint64_t irej[6],
		uint64_t *guard_fired, uint64_t *cap_hit, uint8_t *guard_map);
#define SO_ITER_CAP 4096

/* one pixel column, for unit tests: stack[] is consumed; returns the output
 * sample; rej[2] is incremented like crej[] */
uint16_t so_reject_column(uint16_t *stack, int n, int rejection, double sig_low, double sig_high,
		int sd_mode, uint64_t rej[2], int *guard_fired, int *cap_hit);

/* statistics(fit, 0, NULL, STATS_EXTRA, nullcheck) reduced to what normalization reads
 * (statistics.c:207-326 with IKSS :152-187), literal sort-based form. */
int so_frame_stats_sorted(const uint16_t *frame, int width, int height, int channels,
		double *location, double *scale, double *norm_value, long *ngoodpix);
/* exact 65536-bin histogram of layer 0 (bin 0 = excluded zeros) and the frame max over all layers */
void so_histogram(const uint16_t *frame, int width, int height, int channels,
		uint32_t hist[65536], uint16_t *frame_max);
/* IKSS restated on the histogram (the algorithm the CUDA library runs).
 * norm_value = 65535 or 255 (gui/histogram.c:111-127 via get_normalized_value). */
int so_ikss_hist(const uint32_t hist[65536], double norm_value, double *location, double *scale,
		int *iterations);
/* _compute_normalization_for_image / compute_normalization, stacking.c:79-190.
 * coeff [3][N] = offset, mul, scale. */
int so_norm_coeffs(int mode, int n, int ref_image, const double *location, const double *scale,
		double *coeff);

#ifdef __cplusplus
}
#endif
#endif
