/* Stand-in for <gsl/gsl_statistics_ushort.h> (see ../mini_gsl.c).  Test infrastructure only. */
#ifndef ORACLE_STUB_GSL_STATS_USHORT_H
#define ORACLE_STUB_GSL_STATS_USHORT_H
#include <stddef.h>
double gsl_stats_ushort_mean(const unsigned short data[], const size_t stride, const size_t n);
double gsl_stats_ushort_sd(const unsigned short data[], const size_t stride, const size_t n);
double gsl_stats_ushort_median_from_sorted_data(const unsigned short sorted[], const size_t stride, const size_t n);
double gsl_stats_ushort_absdev_m(const unsigned short data[], const size_t stride, const size_t n, const double mean);
void gsl_stats_ushort_minmax(unsigned short *min, unsigned short *max, const unsigned short data[], const size_t stride, const size_t n);
#endif
