/* Stand-in for <gsl/gsl_histogram.h>; the few calls used are restated in
 * ../mini_gsl.c from GSL's documented behaviour.  Test infrastructure only. */
#ifndef ORACLE_STUB_GSL_HISTOGRAM_H
#define ORACLE_STUB_GSL_HISTOGRAM_H
#include <stddef.h>
typedef struct { size_t n; double *range; double *bin; } gsl_histogram;
gsl_histogram *gsl_histogram_alloc(size_t n);
void gsl_histogram_free(gsl_histogram *h);
int gsl_histogram_set_ranges_uniform(gsl_histogram *h, double xmin, double xmax);
int gsl_histogram_increment(gsl_histogram *h, double x);
double gsl_histogram_get(const gsl_histogram *h, size_t i);
size_t gsl_histogram_bins(const gsl_histogram *h);
#endif
