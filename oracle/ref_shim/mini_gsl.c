/* oracle/ref_shim/mini_gsl.c -- TEST INFRASTRUCTURE ONLY (never linked into the
 * product library).
 *
 * GSL is an un-vendored dependency of the reference (configure.ac:68-73, no
 * pinned version).  The handful of GSL routines the stacking hot path calls
 * are restated here from GSL's published algorithms (statistics/mean_source.c,
 * variance_source.c, median_source.c, fit/linear.c, histogram/*.c as shipped
 * in GSL 1.16 / 2.x -- these routines did not change between those versions).
 *
 * Call sites in the reference:
 *   gsl_stats_ushort_sd                       stacking.c:1676,1698,1713,1727
 *   gsl_stats_ushort_median_from_sorted_data  stacking.c:767,1662,1678,1700,1715,1725
 *   gsl_fit_linear                            stacking.c:1760
 *   gsl_stats_median_from_sorted_data         statistics.c:96,165
 *   gsl_histogram_*                           statistics.c:47-81, gui/histogram.c:111-127
 *
 * long double here is the x87 80-bit type on x86-64, as in a stock GSL build.
 */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>
#include "gsl/gsl_histogram.h"
#include "gsl/gsl_fit.h"
#include "gsl/gsl_statistics.h"

/* running mean in extended precision: m_{i+1} = m_i + (x_i - m_i)/(i+1) */
double gsl_stats_ushort_mean(const unsigned short data[], const size_t stride, const size_t n) {
	long double mean = 0;
	size_t i;
	for (i = 0; i < n; i++)
		mean += (data[i * stride] - mean) / (i + 1);
	return mean;
}

/* running (biased) variance about a double-rounded mean, extended precision */
static double oracle_variance_about(const unsigned short data[], const size_t stride,
		const size_t n, const double mean) {
	long double variance = 0;
	size_t i;
	for (i = 0; i < n; i++) {
		const long double delta = (data[i * stride] - mean);
		variance += (delta * delta - variance) / (i + 1);
	}
	return variance;
}

double gsl_stats_ushort_sd(const unsigned short data[], const size_t stride, const size_t n) {
	const double mean = gsl_stats_ushort_mean(data, stride, n);
	const double variance = oracle_variance_about(data, stride, n, mean);
	return sqrt(variance * ((double) n / (double) (n - 1)));
}

double gsl_stats_ushort_median_from_sorted_data(const unsigned short sorted[],
		const size_t stride, const size_t n) {
	const size_t lhs = (n - 1) / 2;
	const size_t rhs = n / 2;
	if (n == 0)
		return 0.0;
	if (lhs == rhs)
		return sorted[lhs * stride];
	return (sorted[lhs * stride] + sorted[rhs * stride]) / 2.0;
}

double gsl_stats_median_from_sorted_data(const double sorted[], const size_t stride,
		const size_t n) {
	const size_t lhs = (n - 1) / 2;
	const size_t rhs = n / 2;
	if (n == 0)
		return 0.0;
	if (lhs == rhs)
		return sorted[lhs * stride];
	return (sorted[lhs * stride] + sorted[rhs * stride]) / 2.0;
}

double gsl_stats_ushort_absdev_m(const unsigned short data[], const size_t stride,
		const size_t n, const double mean) {
	double sum = 0;
	size_t i;
	for (i = 0; i < n; i++)
		sum += fabs(data[i * stride] - mean);
	return sum / n;
}

void gsl_stats_ushort_minmax(unsigned short *min_out, unsigned short *max_out,
		const unsigned short data[], const size_t stride, const size_t n) {
	unsigned short lo = data[0], hi = data[0];
	size_t i;
	for (i = 0; i < n; i++) {
		unsigned short v = data[i * stride];
		if (v < lo) lo = v;
		if (v > hi) hi = v;
	}
	*min_out = lo;
	*max_out = hi;
}

/* least squares y = c0 + c1 x, running-mean form, all double (fit/linear.c) */
int gsl_fit_linear(const double *x, const size_t xstride, const double *y, const size_t ystride,
		const size_t n, double *c0, double *c1, double *cov_00, double *cov_01,
		double *cov_11, double *sumsq) {
	double m_x = 0, m_y = 0, m_dx2 = 0, m_dxdy = 0;
	size_t i;
	for (i = 0; i < n; i++) {
		m_x += (x[i * xstride] - m_x) / (i + 1.0);
		m_y += (y[i * ystride] - m_y) / (i + 1.0);
	}
	for (i = 0; i < n; i++) {
		const double dx = x[i * xstride] - m_x;
		const double dy = y[i * ystride] - m_y;
		m_dx2 += (dx * dx - m_dx2) / (i + 1.0);
		m_dxdy += (dx * dy - m_dxdy) / (i + 1.0);
	}
	{
		double s2 = 0, d2 = 0;
		double b = m_dxdy / m_dx2;
		double a = m_y - m_x * b;
		*c0 = a;
		*c1 = b;
		for (i = 0; i < n; i++) {
			const double dx = x[i * xstride] - m_x;
			const double dy = y[i * ystride] - m_y;
			const double d = dy - b * dx;
			d2 += d * d;
		}
		s2 = d2 / (n - 2.0);
		*cov_00 = s2 * (1.0 / n) * (1 + m_x * m_x / m_dx2);
		*cov_11 = s2 * 1.0 / (n * m_dx2);
		*cov_01 = s2 * (-m_x) / (n * m_dx2);
		*sumsq = d2;
	}
	return 0;
}

/* ---- histogram: uniform ranges, increment by lookup (histogram/add.c, find.c) ---- */
gsl_histogram *gsl_histogram_alloc(size_t n) {
	gsl_histogram *h = malloc(sizeof(*h));
	h->n = n;
	h->range = malloc((n + 1) * sizeof(double));
	h->bin = calloc(n, sizeof(double));
	return h;
}

void gsl_histogram_free(gsl_histogram *h) {
	if (!h) return;
	free(h->range);
	free(h->bin);
	free(h);
}

int gsl_histogram_set_ranges_uniform(gsl_histogram *h, double xmin, double xmax) {
	size_t i;
	const size_t n = h->n;
	for (i = 0; i <= n; i++) {
		double f1 = ((double) (n - i) / (double) n);
		double f2 = ((double) i / (double) n);
		h->range[i] = f1 * xmin + f2 * xmax;
	}
	for (i = 0; i < n; i++)
		h->bin[i] = 0;
	return 0;
}

/* GSL's find(): linear guess, then bisection; x outside [range[0],range[n]) is rejected */
int gsl_histogram_increment(gsl_histogram *h, double x) {
	const size_t n = h->n;
	const double *range = h->range;
	size_t i_linear, lower, upper, mid;
	if (x < range[0] || x >= range[n])
		return 1;	/* GSL_EDOM */
	{
		double u = (x - range[0]) / (range[n] - range[0]);
		i_linear = (size_t) (u * n);
	}
	if (x >= range[i_linear] && x < range[i_linear + 1]) {
		h->bin[i_linear] += 1.0;
		return 0;
	}
	upper = n;
	lower = 0;
	while (upper - lower > 1) {
		mid = (upper + lower) / 2;
		if (x >= range[mid])
			lower = mid;
		else
			upper = mid;
	}
	h->bin[lower] += 1.0;
	return 0;
}

double gsl_histogram_get(const gsl_histogram *h, size_t i) {
	return h->bin[i];
}

size_t gsl_histogram_bins(const gsl_histogram *h) {
	return h->n;
}
