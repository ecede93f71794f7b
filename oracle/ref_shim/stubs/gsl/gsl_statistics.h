/* Stand-in for <gsl/gsl_statistics.h> (see ../mini_gsl.c).  Test infrastructure only. */
#ifndef ORACLE_STUB_GSL_STATS_H
#define ORACLE_STUB_GSL_STATS_H
#include <stddef.h>
#include "gsl_statistics_ushort.h"
double gsl_stats_median_from_sorted_data(const double sorted[], const size_t stride, const size_t n);
#endif
