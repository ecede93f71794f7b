/* Stand-in for <gsl/gsl_fit.h> (see ../mini_gsl.c).  Test infrastructure only. */
#ifndef ORACLE_STUB_GSL_FIT_H
#define ORACLE_STUB_GSL_FIT_H
#include <stddef.h>
int gsl_fit_linear(const double *x, const size_t xstride, const double *y, const size_t ystride,
		const size_t n, double *c0, double *c1, double *cov00, double *cov01, double *cov11,
		double *sumsq);
#endif
