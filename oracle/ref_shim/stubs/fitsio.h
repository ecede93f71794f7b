/* Stand-in for cfitsio's <fitsio.h>: just the names the two reference files
 * mention.  The oracle feeds frames through a SEQ_SER-typed in-memory
 * sequence, so none of the cfitsio calls is reached at run time (they are
 * defined in ../shim.c and abort if called).  Test infrastructure only. */
#ifndef ORACLE_STUB_FITSIO_H
#define ORACLE_STUB_FITSIO_H
#include <stdio.h>
typedef struct oracle_fitsfile fitsfile;
#define FLEN_FILENAME 1025
#define FLEN_KEYWORD 75
#define FLEN_CARD 81
#define FLEN_VALUE 71
#define FLEN_COMMENT 73
#define FLEN_ERRMSG 81
#define FLEN_STATUS 31
#define BYTE_IMG 8
#define SHORT_IMG 16
#define LONG_IMG 32
#define USHORT_IMG 20
#define FLOAT_IMG -32
#define DOUBLE_IMG -64
#define TBYTE 11
#define TUSHORT 20
#define TSHORT 21
#define TINT 31
#define TLONG 41
#define TFLOAT 42
#define TDOUBLE 82
#define TSTRING 16
int fits_get_img_param(fitsfile *f, int maxaxis, int *bitpix, int *naxis, long *naxes, int *status);
int fits_read_key(fitsfile *f, int datatype, const char *keyname, void *value, char *comm, int *status);
int fits_is_reentrant(void);
void fits_report_error(FILE *stream, int status);
int fits_img_stats_ushort(unsigned short *array, long nx, long ny, int nullcheck,
		unsigned short nullvalue, long *ngoodpix, unsigned short *minvalue,
		unsigned short *maxvalue, double *mean, double *sigma, double *noise1,
		double *noise2, double *noise3, double *noise5, int *status);
#endif
