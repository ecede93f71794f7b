/* Stand-in for <gsl/gsl_matrix.h>: only the type names that reference headers mention. */
#ifndef ORACLE_STUB_GSL_MATRIX_H
#define ORACLE_STUB_GSL_MATRIX_H
#include <stddef.h>
typedef struct { size_t size1, size2, tda; double *data; void *block; int owner; } gsl_matrix;
typedef struct { size_t size, stride; double *data; void *block; int owner; } gsl_vector;
#endif
