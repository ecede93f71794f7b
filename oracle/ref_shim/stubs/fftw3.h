/* Stand-in for <fftw3.h>: type names only. */
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
typedef double fftw_complex[2];
typedef struct oracle_fftw_plan_s *fftw_plan;
#endif
