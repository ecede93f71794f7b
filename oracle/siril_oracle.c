/* oracle/siril_oracle.c -- CPU restatement of Siril 0.9.7's stacking hot path.
 * TEST INFRASTRUCTURE ONLY (see siril_oracle.h for the pinning statement).
 *
 * Written from the semantics of the reference, one function per reference
 * function; every function cites the lines it follows.  Plain C, no GSL:
 * the three GSL routines are restated from GSL's published algorithms
 * (statistics/mean_source.c, variance_source.c, median_source.c, fit/linear.c).
 * Build: oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "siril_oracle.h"

/* ------------------------------------------------------------------------- */
/* primitives                                                                */
/* ------------------------------------------------------------------------- */

/* src/core/utils.c:68-74 */
uint16_t so_round_to_word(double x) {
	/* NaN (inf - inf with the coefficients of a frame whose IKSS scale is 0): the reference's
	 * (WORD)(x + 0.5) is undefined in C; its x86-64 build converts through cvttsd2si, whose
	 * "integer indefinite" 0x80000000 truncates to 0 -- spelled out here */
	if (!(x > 0.0))
		return 0;
	if (x > 65535.0)
		return 65535;
	return (uint16_t) (x + 0.5);
}

/* quicksort_s / quicksort_d (src/core/utils.c:488-533) only matter through their
 * result, the ascending multiset; any correct sort restates them. */
static void sort_u16(uint16_t *a, int n) {
	int i, j;
	if (n <= 48) {
		for (i = 1; i < n; i++) {
			uint16_t v = a[i];
			for (j = i; j > 0 && a[j - 1] > v; j--)
				a[j] = a[j - 1];
			a[j] = v;
		}
		return;
	}
	{	/* LSD radix, two 8-bit passes */
		uint16_t *tmp = malloc((size_t) n * sizeof(uint16_t));
		int cnt[256], pass;
		uint16_t *src = a, *dst = tmp;
		for (pass = 0; pass < 2; pass++) {
			int shift = pass * 8, acc = 0;
			memset(cnt, 0, sizeof(cnt));
			for (i = 0; i < n; i++)
				cnt[(src[i] >> shift) & 255]++;
			for (i = 0; i < 256; i++) {
				int c = cnt[i];
				cnt[i] = acc;
				acc += c;
			}
			for (i = 0; i < n; i++)
				dst[cnt[(src[i] >> shift) & 255]++] = src[i];
			{ uint16_t *t = src; src = dst; dst = t; }
		}
		free(tmp);	/* two passes: the result is back in a[] */
	}
}

static int cmp_double(const void *pa, const void *pb) {
	double a = *(const double *) pa, b = *(const double *) pb;
	return (a > b) - (a < b);
}

/* gsl_stats_ushort_median_from_sorted_data (GSL statistics/median_source.c) */
static double median_sorted_u16(const uint16_t *s, int n) {
	int lhs = (n - 1) / 2, rhs = n / 2;
	if (n == 0)
		return 0.0;
	if (lhs == rhs)
		return s[lhs];
	return (s[lhs] + s[rhs]) / 2.0;
}

/* gsl_stats_median_from_sorted_data */
static double median_sorted_f64(const double *s, size_t n) {
	size_t lhs = (n - 1) / 2, rhs = n / 2;
	if (n == 0)
		return 0.0;
	if (lhs == rhs)
		return s[lhs];
	return (s[lhs] + s[rhs]) / 2.0;
}

/* gsl_stats_ushort_sd (GSL statistics/variance_source.c + mean_source.c):
 * running mean and running biased variance in long double, mean rounded to
 * double in between, unbiased by n/(n-1) in double.  Call sites
 * stacking.c:1676,1698,1713,1727. */
static double sd_gsl(const uint16_t *d, int n) {
	long double mean_ld = 0, var = 0;
	double mean;
	int i;
	for (i = 0; i < n; i++)
		mean_ld += (d[i] - mean_ld) / (i + 1);
	mean = mean_ld;
	for (i = 0; i < n; i++) {
		const long double delta = (d[i] - mean);
		var += (delta * delta - var) / (i + 1);
	}
	return sqrt((double) var * ((double) n / (double) (n - 1)));
}

/* the same quantity from exact integer sums: n*Q - S*S is an exact integer
 * (< 2^64 for n < 65537), so the biased variance carries one rounding */
static double sd_exact(const uint16_t *d, int n) {
	uint64_t S = 0, Q = 0, D;
	double var;
	int i;
	for (i = 0; i < n; i++) {
		S += d[i];
		Q += (uint64_t) d[i] * d[i];
	}
	D = (uint64_t) n * Q - S * S;
	var = (double) D / (double) ((uint64_t) n * (uint64_t) n);
	return sqrt(var * ((double) n / (double) (n - 1)));
}

static double sd_u16(const uint16_t *d, int n, int sd_mode) {
	return (sd_mode & SO_SD_EXACT_INTEGER) ? sd_exact(d, n) : sd_gsl(d, n);
}

/* gsl_fit_linear (GSL fit/linear.c) with x_i = i: intercept c0, slope c1.
 * Call site stacking.c:1756-1760. */
static void fit_line_index(const uint16_t *y, int n, double *c0, double *c1) {
	double m_x = 0, m_y = 0, m_dx2 = 0, m_dxdy = 0;
	int i;
	for (i = 0; i < n; i++) {
		m_x += ((double) i - m_x) / (i + 1.0);
		m_y += ((double) y[i] - m_y) / (i + 1.0);
	}
	for (i = 0; i < n; i++) {
		const double dx = (double) i - m_x;
		const double dy = (double) y[i] - m_y;
		m_dx2 += (dx * dx - m_dx2) / (i + 1.0);
		m_dxdy += (dx * dy - m_dxdy) / (i + 1.0);
	}
	*c1 = m_dxdy / m_dx2;
	*c0 = m_y - m_x * (*c1);
}

/* The same least-squares line and mean absolute residual from exact sums (SO_FIT_EXACT_SUMS; the
 * GPU's form, csrc/ssc_fit.cuh).  With x_i = i the normal equations are integer:
 *   slope a     = (n*Sxy - Sx*Sy) / (n*Sxx - Sx*Sx),  Sx = n(n-1)/2,  n*Sxx - Sx^2 = n^2 (n^2 - 1) / 12
 *   intercept b = Sy/n - ((n-1)/2) * a
 * numerator and denominator are exact integers below 2^53 (n <= 2048): ONE rounding for a, three for b,
 * against the reference's 4n sequentially rounded running-mean updates.  The residual of sample f is
 * y_f - fma(a, f, b); |residual| is rounded to a multiple of 2^-32 and summed as an integer, so the
 * sum does not depend on the order (any thread layout gives the same bits);
 * sigma = (that sum * 2^-32) / n.  Flags compare the same residual with slow*sigma and shigh*sigma
 * (the reference divides the residual by sigma instead, stacking.c:1175,1179). */
static void fit_exact(const uint16_t *y, int n, double *c0, double *c1, double *sigma) {
	uint64_t Sy = 0, Sxy = 0, T = 0;
	int i;
	for (i = 0; i < n; i++) {
		Sy += y[i];
		Sxy += (uint64_t) i * y[i];
	}
	{
		const int64_t num = (int64_t) ((uint64_t) n * Sxy) - (int64_t) (((uint64_t) n * (uint64_t) (n - 1) / 2) * Sy);
		const uint64_t n2 = (uint64_t) n * (uint64_t) n;
		const uint64_t den = n2 * (n2 - 1) / 12;
		const double a = (double) num / (double) den;
		*c1 = a;
		*c0 = (double) Sy / (double) n - ((double) (n - 1) * 0.5) * a;
	}
	for (i = 0; i < n; i++) {
		const double res = (double) y[i] - fma(*c1, (double) i, *c0);
		T += (uint64_t) llrint(fabs(res) * 4294967296.0);	/* round to nearest even, as the GPU's 2^52 trick */
	}
	*sigma = ((double) T * (1.0 / 4294967296.0)) / (double) n;
}

/* ------------------------------------------------------------------------- */
/* clipping predicates, stacking.c:1130-1183                                 */
/* ------------------------------------------------------------------------- */

/* percentile_clipping, stacking.c:1130-1143 */
static int clip_percentile(uint16_t p, double plow, double phigh, double median, uint64_t rej[2]) {
	if ((median - (double) p) / median > plow) {
		rej[0]++;
		return -1;
	}
	if (((double) p - median) / median > phigh) {
		rej[1]++;
		return 1;
	}
	return 0;
}

/* sigma_clipping, stacking.c:1148-1161 */
static int clip_sigma(uint16_t p, double slow, double shigh, double sigma, double median, uint64_t rej[2]) {
	if (median - (double) p > slow * sigma) {
		rej[0]++;
		return -1;
	}
	if ((double) p - median > shigh * sigma) {
		rej[1]++;
		return 1;
	}
	return 0;
}

/* line_clipping, stacking.c:1170-1183 */
static int clip_line(uint16_t p, double slow, double shigh, double sigma, int i, double a, double b,
		uint64_t rej[2]) {
	if (((a * (double) i + b - (double) p) / sigma) > slow) {
		rej[0]++;
		return -1;
	}
	if ((((double) p - a * (double) i - b) / sigma) > shigh) {
		rej[1]++;
		return 1;
	}
	return 0;
}

/* drop the flagged samples keeping order (remove_pixel loops, stacking.c:1686-1692) */
static int compact_flagged(uint16_t *s, const signed char *flag, int n) {
	int j, out = 0;
	for (j = 0; j < n; j++)
		if (!flag[j])
			s[out++] = s[j];
	return n - out;
}

/* One pixel column through stacking.c:1656-1794.
 *
 * The one deliberate difference: when the reference leaves its flagging loop
 * through "if (N - r <= 4) break;" (stacking.c:1684,1737,1771) the flags of the
 * samples it did not reach keep whatever an earlier iteration or an earlier
 * pixel of the same OpenMP thread wrote there (r is never reset either).  That
 * is history dependent; here those flags read as 0 and the event is counted in
 * *guard_fired so that parity runs can assert it never happened. */
uint16_t so_reject_column(uint16_t *s, int N, int rejection, double sig_low, double sig_high,
		int sd_mode, uint64_t rej[2], int *guard_fired, int *cap_hit) {
	signed char flag_small[256];
	signed char *flag = N <= 256 ? flag_small : malloc((size_t) N);
	uint16_t *w = NULL;
	double sigma, median, sum;
	int n, f, r = 0, iter = 0;

	switch (rejection) {
	case SO_PERCENTILE:	/* stacking.c:1660-1673 */
		sort_u16(s, N);
		median = median_sorted_u16(s, N);
		for (f = 0; f < N; f++)
			flag[f] = (signed char) clip_percentile(s[f], sig_low, sig_high, median, rej);
		{
			int cur = N, out = 0;
			for (f = 0; f < N; f++) {
				if (flag[f] != 0 && cur > 1)
					cur--;
				else
					s[out++] = s[f];
			}
			N = cur;
		}
		break;
	case SO_SIGMA:		/* stacking.c:1674-1695 */
		do {
			sigma = sd_u16(s, N, sd_mode);
			sort_u16(s, N);
			median = median_sorted_u16(s, N);
			memset(flag, 0, (size_t) N);
			for (f = 0; f < N; f++) {
				flag[f] = (signed char) clip_sigma(s[f], sig_low, sig_high, sigma, median, rej);
				if (flag[f])
					r++;
				if (N - r <= 4) {
					if (f < N - 1)
						*guard_fired = 1;
					break;
				}
			}
			n = compact_flagged(s, flag, N);
			N -= n;
			if (++iter >= SO_ITER_CAP) { *cap_hit = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case SO_SIGMEDIAN:	/* stacking.c:1696-1709 */
		do {
			sigma = sd_u16(s, N, sd_mode);
			sort_u16(s, N);
			median = median_sorted_u16(s, N);
			n = 0;
			for (f = 0; f < N; f++) {
				if (clip_sigma(s[f], sig_low, sig_high, sigma, median, rej)) {
					s[f] = so_round_to_word(median);
					n++;
				}
			}
			if (++iter >= SO_ITER_CAP) { *cap_hit = 1; break; }
		} while (n > 0 && N > 3);
		break;
	case SO_WINSORIZED:	/* stacking.c:1710-1749 */
		w = malloc((size_t) N * sizeof(uint16_t));
		do {
			double sigma0;
			int inner = 0;
			sigma = sd_u16(s, N, sd_mode);
			sort_u16(s, N);
			median = median_sorted_u16(s, N);
			memcpy(w, s, (size_t) N * sizeof(uint16_t));
			do {
				double m0 = median - 1.5 * sigma;
				double m1 = median + 1.5 * sigma;
				for (f = 0; f < N; f++) {	/* Winsorized(), stacking.c:1163-1168 */
					if (w[f] < m0)
						w[f] = so_round_to_word(m0);
					else if (w[f] > m1)
						w[f] = so_round_to_word(m1);
				}
				sort_u16(w, N);
				median = median_sorted_u16(w, N);
				sigma0 = sigma;
				sigma = 1.134 * sd_u16(w, N, sd_mode);
				if (++inner >= SO_ITER_CAP) { *cap_hit = 1; break; }
			} while ((fabs(sigma - sigma0) / sigma0) > 0.0005);
			memset(flag, 0, (size_t) N);
			for (f = 0; f < N; f++) {
				flag[f] = (signed char) clip_sigma(s[f], sig_low, sig_high, sigma, median, rej);
				if (flag[f] != 0)
					r++;
				if (N - r <= 4) {
					if (f < N - 1)
						*guard_fired = 1;
					break;
				}
			}
			n = compact_flagged(s, flag, N);
			N -= n;
			if (++iter >= SO_ITER_CAP) { *cap_hit = 1; break; }
		} while (n > 0 && N > 3);
		free(w);
		break;
	case SO_LINEARFIT:	/* stacking.c:1750-1784 */
		do {
			double a, b;
			const int exact = (sd_mode & SO_FIT_EXACT_SUMS) != 0;
			double tl = 0.0, th = 0.0;
			sort_u16(s, N);
			if (exact) {
				fit_exact(s, N, &b, &a, &sigma);
				tl = sig_low * sigma;
				th = sig_high * sigma;
			} else {
				fit_line_index(s, N, &b, &a);
				sigma = 0.0;
				for (f = 0; f < N; f++)
					sigma += fabs((double) s[f] - (a * (double) f + b));
				sigma /= (double) N;
			}
			memset(flag, 0, (size_t) N);
			for (f = 0; f < N; f++) {
				if (exact) {
					const double res = (double) s[f] - fma(a, (double) f, b);
					flag[f] = (signed char) (-res > tl ? -1 : (res > th ? 1 : 0));
					if (flag[f])
						rej[flag[f] < 0 ? 0 : 1]++;
				} else
					flag[f] = (signed char) clip_line(s[f], sig_low, sig_high, sigma, f, a, b, rej);
				if (flag[f] != 0)
					r++;
				if (N - r <= 4) {
					if (f < N - 1)
						*guard_fired = 1;
					break;
				}
			}
			n = compact_flagged(s, flag, N);
			N -= n;
			if (++iter >= SO_ITER_CAP) { *cap_hit = 1; break; }
		} while (n > 0 && N > 3);
		break;
	default:
		break;		/* NO_REJEC */
	}

	sum = 0.0;		/* stacking.c:1790-1794 */
	for (f = 0; f < N; f++)
		sum += s[f];
	if (flag != flag_small)
		free(flag);
	return so_round_to_word(sum / (double) N);
}

/* ------------------------------------------------------------------------- */
/* the three operators                                                       */
/* ------------------------------------------------------------------------- */

static const uint16_t *plane_of(const so_stack_args *a, int frame, int layer) {
	size_t plane = (size_t) a->width * a->height;
	return a->frames + ((size_t) frame * a->channels + layer) * plane;
}

static int use_threads(const so_stack_args *a) {
#ifdef _OPENMP
	return a->threads > 0 ? a->threads : omp_get_num_procs();
#else
	(void) a;
	return 1;
#endif
}

/* stack_summing, stacking.c:196-355: sums in frame order with the registration
 * shift (:299-319, source index 0 is never read, :307), global max over all
 * layers, rescale by 65535/max when it does not fit 16 bits (:328-341).
 * Frames arrive through seq_read_frame, i.e. vertically flipped for SER. */
int so_stack_sum(const so_stack_args *a, uint16_t *out, uint64_t *maxim_out) {
	const int W = a->width, H = a->height, C = a->channels, N = a->n_frames;
	const size_t plane = (size_t) W * H;
	uint64_t *acc, maxim = 0;
	double ratio;
	int j, layer, x, y;
	size_t i;
	if (N <= 1)
		return -1;
	acc = calloc(plane * C, sizeof(uint64_t));
	if (!acc)
		return -2;
	for (j = 0; j < N; j++) {
		int sx = a->shiftx ? a->shiftx[j] : 0, sy = a->shifty ? a->shifty[j] : 0;
		for (layer = 0; layer < C; layer++) {
			const uint16_t *src = plane_of(a, j, layer);
			uint64_t *dst = acc + layer * plane;
			for (y = 0; y < H; y++) {
				int ny = y - sy;
				if (ny < 0 || ny >= H)
					continue;
				for (x = 0; x < W; x++) {
					int nx = x - sx;
					size_t ii;
					if (nx < 0 || nx >= W)
						continue;
					ii = (size_t) ny * W + nx;
					if (ii == 0)
						continue;
					{
						int srow = a->row_order ? H - 1 - ny : ny;
						uint64_t v = (dst[(size_t) y * W + x] += src[(size_t) srow * W + nx]);
						if (v > maxim)
							maxim = v;
					}
				}
			}
		}
	}
	ratio = maxim > 65535 ? 65535.0 / (double) maxim : 1.0;
	for (i = 0; i < plane * C; i++)
		out[i] = ratio == 1.0 ? so_round_to_word((double) acc[i]) : so_round_to_word((double) acc[i] * ratio);
	if (maxim_out)
		*maxim_out = maxim;
	free(acc);
	return 0;
}

/* stack_addmax / stack_addmin: the brightest / darkest sample per pixel over the frames, with the
 * registration shifts and the "ii > 0" test of stack_summing (stacking.c:918-936, 1074-1092).
 * Pixels no frame reaches keep the initial value: 0 (calloc, :882) or 0xFFFF (memset, :1037-1038). */
int so_stack_extreme(const so_stack_args *a, int take_max, uint16_t *out, uint16_t *extreme_out) {
	const int W = a->width, H = a->height, C = a->channels, N = a->n_frames;
	const size_t plane = (size_t) W * H;
	uint16_t extreme = take_max ? 0 : 0xFFFF;
	int j, layer, x, y;
	size_t i;
	if (N <= 1)
		return -1;
	for (i = 0; i < plane * C; i++)
		out[i] = take_max ? 0 : 0xFFFF;
	for (j = 0; j < N; j++) {
		int sx = a->shiftx ? a->shiftx[j] : 0, sy = a->shifty ? a->shifty[j] : 0;
		for (layer = 0; layer < C; layer++) {
			const uint16_t *src = plane_of(a, j, layer);
			uint16_t *dst = out + layer * plane;
			for (y = 0; y < H; y++) {
				int ny = y - sy, srow;
				if (ny < 0 || ny >= H)
					continue;
				srow = a->row_order ? H - 1 - ny : ny;
				for (x = 0; x < W; x++) {
					int nx = x - sx;
					uint16_t v, *d = dst + (size_t) y * W + x;
					if (nx < 0 || nx >= W || (ny == 0 && nx == 0))
						continue;
					v = src[(size_t) srow * W + nx];
					if (take_max ? v > *d : v < *d)
						*d = v;
					if (take_max ? *d > extreme : *d < extreme)
						extreme = *d;
				}
			}
		}
	}
	if (extreme_out)
		*extreme_out = extreme;
	return 0;
}

/* normalization of one sample on the median path, stacking.c:750-764: always
 * through double, NO_NORM included */
static uint16_t normalize_median_path(uint16_t p, int mode, double offset, double mul, double scale) {
	double tmp = (double) p * scale;
	if (mode == SO_MULTIPLICATIVE || mode == SO_MULTIPLICATIVE_SCALING)
		return so_round_to_word(tmp * mul);
	return so_round_to_word(tmp - offset);
}

/* stack_median, stacking.c:362-816.  Registration shifts are ignored by the
 * reference on this path.  Block row t (top-down) is written to storage row
 * H-t-1 (:731); the feed hands out row t from storage row H-1-t (FITS) or t (SER). */
int so_stack_median(const so_stack_args *a, uint16_t *out) {
	const int W = a->width, H = a->height, C = a->channels, N = a->n_frames;
	const size_t plane = (size_t) W * H;
	int row;
	if (N < 2)
		return -1;
#pragma omp parallel for schedule(static) num_threads(use_threads(a))
	for (row = 0; row < C * H; row++) {
		int layer = row / H, t = row % H, x, f;
		int srow = a->row_order ? t : H - 1 - t;
		uint16_t *stack = malloc((size_t) N * sizeof(uint16_t));
		uint16_t *dst = out + layer * plane + (size_t) (H - 1 - t) * W;
		for (x = 0; x < W; x++) {
			for (f = 0; f < N; f++) {
				uint16_t p = plane_of(a, f, layer)[(size_t) srow * W + x];
				double off = a->coeff ? a->coeff[f] : 0.0;
				double mul = a->coeff ? a->coeff[N + f] : 1.0;
				double sc = a->coeff ? a->coeff[2 * N + f] : 1.0;
				stack[f] = normalize_median_path(p, a->normalize, off, mul, sc);
			}
			sort_u16(stack, N);
			dst[x] = (uint16_t) median_sorted_u16(stack, N);	/* double -> WORD truncation, :766-767 */
		}
		free(stack);
	}
	return 0;
}

/* stack_mean_with_rejection, stacking.c:1189-1858.
 * Gather (:1535-1591 y shift by moving the read area, rows outside the image are
 * zero-filled BEFORE normalization; :1623-1654 x shift, columns outside the image
 * give a literal 0 AFTER normalization), normalize (:1635-1652, NO_NORM copies the
 * raw sample), reject, mean, per-channel counters (:1796-1802).
 * The per-row formulation below equals the reference's per-block one whenever
 * the reference's own block arithmetic is well defined (its "offset" for a block
 * that starts below row 0 after the shift, :1560, is only right for the first
 * block of a channel). */
int so_stack_mean_map(const so_stack_args *a, uint16_t *out, uint64_t irej[6],
		uint64_t *guard_fired, uint64_t *cap_hit, uint8_t *guard_map) {
	const int W = a->width, H = a->height, C = a->channels, N = a->n_frames;
	const size_t plane = (size_t) W * H;
	const int have_reg = a->shiftx || a->shifty;
	uint64_t rej_tot[6] = { 0, 0, 0, 0, 0, 0 }, guards = 0, caps = 0;
	int row;
	if (N < 2)
		return -1;
#pragma omp parallel for schedule(static) num_threads(use_threads(a))
	for (row = 0; row < C * H; row++) {
		int layer = row / H, t = row % H, x, f;
		uint16_t *stack = malloc((size_t) N * sizeof(uint16_t));
		uint16_t *dst = out + layer * plane + (size_t) (H - 1 - t) * W;
		uint64_t crej[2] = { 0, 0 }, g = 0, c = 0;
		for (x = 0; x < W; x++) {
			int gf = 0, ch = 0;
			for (f = 0; f < N; f++) {
				int sx = (have_reg && a->shiftx) ? a->shiftx[f] : 0;
				int sy = (have_reg && a->shifty) ? a->shifty[f] : 0;
				int ty = t + sy;	/* top-down source row */
				uint16_t p;
				if (sx && (x - sx >= W || x - sx < 0)) {
					stack[f] = 0;
					continue;
				}
				if (ty < 0 || ty >= H) {
					p = 0;
				} else {
					int srow = a->row_order ? ty : H - 1 - ty;
					p = plane_of(a, f, layer)[(size_t) srow * W + (x - sx)];
				}
				switch (a->normalize) {
				default:
					stack[f] = p;
					break;
				case SO_ADDITIVE:
				case SO_ADDITIVE_SCALING: {
					double tmp = (double) p * (a->coeff ? a->coeff[2 * N + f] : 1.0);
					stack[f] = so_round_to_word(tmp - (a->coeff ? a->coeff[f] : 0.0));
					break;
				}
				case SO_MULTIPLICATIVE:
				case SO_MULTIPLICATIVE_SCALING: {
					double tmp = (double) p * (a->coeff ? a->coeff[2 * N + f] : 1.0);
					stack[f] = so_round_to_word(tmp * (a->coeff ? a->coeff[N + f] : 1.0));
					break;
				}
				}
			}
			dst[x] = so_reject_column(stack, N, a->rejection, a->sig_low, a->sig_high,
					a->sd_mode, crej, &gf, &ch);
			g += gf;
			c += ch;
			if (guard_map)
				guard_map[layer * plane + (size_t) (H - 1 - t) * W + x] = (uint8_t) gf;
		}
		free(stack);
#pragma omp critical
		{
			rej_tot[layer * 2 + 0] += crej[0];
			rej_tot[layer * 2 + 1] += crej[1];
			guards += g;
			caps += c;
		}
	}
	if (irej)
		memcpy(irej, rej_tot, sizeof(rej_tot));
	if (guard_fired)
		*guard_fired = guards;
	if (cap_hit)
		*cap_hit = caps;
	return 0;
}

int so_stack_mean(const so_stack_args *a, uint16_t *out, uint64_t irej[6],
		uint64_t *guard_fired, uint64_t *cap_hit) {
	return so_stack_mean_map(a, out, irej, guard_fired, cap_hit, NULL);
}

/* ------------------------------------------------------------------------- */
/* normalization statistics                                                  */
/* ------------------------------------------------------------------------- */

/* siril_stats_double_mad, statistics.c:83-100 */
static double mad_f64(const double *d, size_t n, double m) {
	double *tmp = malloc(n * sizeof(double)), r;
	size_t i;
	for (i = 0; i < n; i++)
		tmp[i] = fabs(d[i] - m);
	qsort(tmp, n, sizeof(double), cmp_double);
	r = median_sorted_f64(tmp, n);
	free(tmp);
	return r;
}

/* one term of the biweight midvariance sums, statistics.c:136-145 */
static void bwmv_term(double d, double m, double mad, double *up, double *down) {
	double yi = (d - m) / (9 * mad);
	double yi2 = yi * yi;
	double ai = (fabs(yi) < 1.0) ? 1.0 : 0.0;
	*up = ai * ((d - m) * (d - m)) * (((1 - yi2) * (1 - yi2)) * ((1 - yi2) * (1 - yi2)));
	*down = (ai * (1 - yi2) * (1 - 5 * yi2));
}

/* siril_stats_double_bwmv, statistics.c:128-150 */
static double bwmv_f64(const double *d, size_t n, double mad, double m) {
	double up = 0.0, down = 0.0, u, w;
	size_t i;
	if (!(mad > 0.0))
		return 0.0;
	for (i = 0; i < n; i++) {
		bwmv_term(d[i], m, mad, &u, &w);
		up += u;
		down += w;
	}
	return n * (up / (down * down));
}

/* statistics() as normalization calls it (sequence.c:1110): layer 0, zeros
 * excluded, IKSS (statistics.c:152-187) on data/(hist_size-1), scaled back
 * (:279-288).  hist_size-1 is 255 when the frame's max over ALL layers is
 * <= 255 (gui/histogram.c:116, utils.c:454-459, single_image.c:220-250). */
int so_frame_stats_sorted(const uint16_t *frame, int width, int height, int channels,
		double *location, double *scale, double *norm_value, long *ngoodpix) {
	size_t plane = (size_t) width * height, i, n = 0, lo, hi;
	uint16_t mx = 0;
	double norm, *d, s0 = 1, m = 0, s = 0;
	for (i = 0; i < plane * channels; i++)
		if (frame[i] > mx)
			mx = frame[i];
	norm = mx <= 255 ? 255.0 : 65535.0;
	for (i = 0; i < plane; i++)
		if (frame[i] > 0)
			n++;
	*norm_value = norm;
	*ngoodpix = (long) n;
	if (n == 0)
		return -1;	/* statistics() returns NULL, :247-250 */
	d = malloc(n * sizeof(double));
	for (i = 0, lo = 0; i < plane; i++)
		if (frame[i] > 0)
			d[lo++] = (double) frame[i] / norm;
	qsort(d, n, sizeof(double), cmp_double);
	lo = 0;
	hi = n;
	for (;;) {
		double mad, xlow, xhigh;
		if (hi - lo < 1) {
			*location = *scale = 0;
			break;
		}
		m = median_sorted_f64(d + lo, hi - lo);
		mad = mad_f64(d + lo, hi - lo, m);
		s = sqrt(bwmv_f64(d + lo, hi - lo, mad, m));
		if (s < 2E-23) {
			*location = m;
			*scale = 0;
			break;
		}
		if (((s0 - s) / s) < 10E-6) {
			*location = m;
			*scale = 0.991 * s;
			break;
		}
		s0 = s;
		xlow = m - 4 * s;
		xhigh = m + 4 * s;
		while (d[lo] < xlow)
			lo++;
		while (d[hi - 1] > xhigh)
			hi--;
	}
	*location *= norm;
	*scale *= norm;
	free(d);
	return 0;
}

void so_histogram(const uint16_t *frame, int width, int height, int channels,
		uint32_t hist[65536], uint16_t *frame_max) {
	size_t plane = (size_t) width * height, i;
	uint16_t mx = 0;
	memset(hist, 0, 65536 * sizeof(uint32_t));
	for (i = 0; i < plane; i++)
		hist[frame[i]]++;
	for (i = 0; i < plane * channels; i++)
		if (frame[i] > mx)
			mx = frame[i];
	*frame_max = mx;
}

/* rank -> bin through the cumulative counts of the window [bi, bj) */
static int bin_of_rank(const uint32_t *hist, int bi, int bj, uint64_t rank) {
	uint64_t acc = 0;
	int k;
	for (k = bi; k < bj; k++) {
		acc += hist[k];
		if (acc > rank)
			return k;
	}
	return bj - 1;
}

/* The rank-th smallest of { |d_k - m| repeated hist[k] times } by a two-sided
 * walk outward from m (deltas are monotone on each side). */
static double mad_delta_of_rank(const uint32_t *hist, int bi, int bj, double norm, double m,
		uint64_t rank) {
	int kl, kr;
	uint64_t acc = 0;
	/* kr = first bin with d_k > m ; kl = kr - 1 */
	for (kr = bi; kr < bj; kr++)
		if ((double) kr / norm > m)
			break;
	kl = kr - 1;
	for (;;) {
		int have_l = kl >= bi, have_r = kr < bj;
		double dl = have_l ? fabs((double) kl / norm - m) : 0, dr = have_r ? fabs((double) kr / norm - m) : 0;
		if (!have_l && !have_r)
			return 0.0;
		if (have_l && (!have_r || dl <= dr)) {
			acc += hist[kl];
			if (acc > rank)
				return dl;
			kl--;
		} else {
			acc += hist[kr];
			if (acc > rank)
				return dr;
			kr++;
		}
	}
}

/* IKSS (statistics.c:152-187) restated on the exact histogram of the non-zero
 * samples.  Every clipping window of the sorted-sample form is a run of whole
 * bins [bi, bj), so medians and MADs are the same order statistics; the BWMV
 * sums add count*term per bin in a fixed order instead of one term per sample:
 * bins in ascending order inside 1024 chunks of 64 bins, then a pairwise tree
 * over the chunk partials (stride 512, 256, ... 1) -- the order the CUDA kernel
 * uses, so both give bit-identical results. */
int so_ikss_hist(const uint32_t hist[65536], double norm, double *location, double *scale,
		int *iterations) {
	int bi = 1, bj = 65536, k, it = 0;
	double s0 = 1, m, s;
	static __thread double pu[1024], pd[1024];
	for (;;) {
		uint64_t n = 0, lhs, rhs;
		double mad, dl, dr, up, down, bw, xlow, xhigh;
		int kl, kr, c, stride;
		it++;
		for (k = bi; k < bj; k++)
			n += hist[k];
		if (n < 1) {
			*location = *scale = 0;
			break;
		}
		lhs = (n - 1) / 2;
		rhs = n / 2;
		kl = bin_of_rank(hist, bi, bj, lhs);
		kr = bin_of_rank(hist, bi, bj, rhs);
		m = lhs == rhs ? (double) kl / norm : ((double) kl / norm + (double) kr / norm) / 2.0;
		dl = mad_delta_of_rank(hist, bi, bj, norm, m, lhs);
		dr = mad_delta_of_rank(hist, bi, bj, norm, m, rhs);
		mad = lhs == rhs ? dl : (dl + dr) / 2.0;
		bw = 0.0;
		if (mad > 0.0) {
			for (c = 0; c < 1024; c++) {
				double u = 0.0, w = 0.0, tu, tw;
				/* partial sum c takes bins bi + c, bi + c + 1024, ... of the clipping window (the
				 * order k_ikss uses: the window's bins spread evenly over its 1024 threads) */
				for (k = bi + c; k < bj; k += 1024) {
					if (hist[k] == 0)
						continue;
					bwmv_term((double) k / norm, m, mad, &tu, &tw);
					u += (double) hist[k] * tu;
					w += (double) hist[k] * tw;
				}
				pu[c] = u;
				pd[c] = w;
			}
			for (stride = 512; stride >= 1; stride >>= 1)
				for (c = 0; c < stride; c++) {
					pu[c] += pu[c + stride];
					pd[c] += pd[c + stride];
				}
			up = pu[0];
			down = pd[0];
			bw = (double) n * (up / (down * down));
		}
		s = sqrt(bw);
		if (s < 2E-23) {
			*location = m;
			*scale = 0;
			break;
		}
		if (((s0 - s) / s) < 10E-6) {
			*location = m;
			*scale = 0.991 * s;
			break;
		}
		s0 = s;
		xlow = m - 4 * s;
		xhigh = m + 4 * s;
		while (bi < bj && (hist[bi] == 0 || (double) bi / norm < xlow))
			bi++;
		while (bj > bi && (hist[bj - 1] == 0 || (double) (bj - 1) / norm > xhigh))
			bj--;
		if (it >= SO_ITER_CAP)
			break;
	}
	*location *= norm;
	*scale *= norm;
	if (iterations)
		*iterations = it;
	return 0;
}

/* compute_normalization + _compute_normalization_for_image, stacking.c:79-190.
 * coeff = [offset[N], mul[N], scale[N]] */
int so_norm_coeffs(int mode, int n, int ref, const double *location, const double *scale_in,
		double *coeff) {
	double *offset = coeff, *mul = coeff + n, *scale = coeff + 2 * n;
	double scale0, mul0, offset0;
	int i;
	for (i = 0; i < n; i++) {
		offset[i] = 0.0;
		mul[i] = 1.0;
		scale[i] = 1.0;
	}
	if (mode == SO_NO_NORM)
		return 0;
	if (ref < 0)
		ref = 0;	/* :143-145 */
	scale0 = scale_in[ref];
	offset0 = mul0 = location[ref];
	for (i = 0; i < n; i++) {
		switch (mode) {
		default:
		case SO_ADDITIVE_SCALING:
			scale[i] = scale0 / scale_in[i];
			offset[i] = scale[i] * location[i] - offset0;
			break;
		case SO_ADDITIVE:
			offset[i] = scale[i] * location[i] - offset0;
			break;
		case SO_MULTIPLICATIVE_SCALING:
			scale[i] = scale0 / scale_in[i];
			mul[i] = mul0 / location[i];
			break;
		case SO_MULTIPLICATIVE:
			mul[i] = mul0 / location[i];
			break;
		}
	}
	return 0;
}
