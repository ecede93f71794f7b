// ssc_device.cuh -- device-side primitives shared by the stacking kernels.
//
// Every floating-point expression that mirrors the reference's double
// arithmetic is spelled with the round-to-nearest intrinsics (__dadd_rn,
// __dmul_rn, __ddiv_rn, __dsqrt_rn): they are never contracted into FMAs, so
// each operation is rounded individually exactly as the reference's x86-64
// build (gcc -O2, no -mfma, configure.ac:38-48) rounds it.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/siril_stack_cuda.h"

namespace ssc {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }

// round_to_WORD, src/core/utils.c:68-74
// NaN (inf - inf with the coefficients of a frame whose IKSS scale is 0) is undefined in the C source;
// the reference's x86-64 build yields 0 (cvttsd2si's 0x80000000 truncated to 16 bits), and so does this.
__device__ __forceinline__ unsigned int round_to_word(double x) {
	if (!(x > 0.0)) return 0u;
	if (x > 65535.0) return 65535u;
	return (unsigned int) __double2int_rz(dadd(x, 0.5));
}

// Normalization of one raw sample.
// mean path: stacking.c:1635-1652 (NO_NORM copies the sample);
// median path: stacking.c:750-764 (always through double; identical results).
__device__ __forceinline__ unsigned int normalize_sample(unsigned int p, int mode, double offset,
		double mul, double scale) {
	switch (mode) {
	case ADDITIVE:
	case ADDITIVE_SCALING:
		return round_to_word(dsub(dmul((double) (int) p, scale), offset));
	case MULTIPLICATIVE:
	case MULTIPLICATIVE_SCALING:
		return round_to_word(dmul(dmul((double) (int) p, scale), mul));
	default:
		return p;
	}
}

// gsl_stats_ushort_sd (stacking.c:1676,1698,1713,1727) from exact integer sums:
// n*Q - S*S is an exact integer below 2^64 for n <= 65536, so the biased
// variance carries a single rounding; then the reference's own
// sqrt(variance * (n / (n-1))).  GSL accumulates the same quantity in x87
// long double, which FP64 cannot reproduce bit for bit.
// Small-n tables (n <= 64) that take the divisions out of the per-pixel loops without changing a
// bit: unbias[n] = n / (n-1) is the reference's own quotient; inv_n[n] = RN(1/n) and
// inv_n2[n] = RN(1/n^2) feed div_exact() below.
struct SmallN {
	const double *unbias, *inv_n, *inv_n2;
};
#define SSC_SMALLN_MAX 64

// a / b correctly rounded, given y = RN(1 / b): q0 = RN(a*y), r = a - b*q0 (exact in one FMA),
// q = RN(q0 + r*y).  Markstein's theorem: this is RN(a / b) whenever y is the correctly rounded
// reciprocal and b's significand is not all ones -- b is a small integer here.  The explicit FMAs
// are part of the division algorithm (as inside __ddiv_rn itself), not contractions of the
// reference's arithmetic.
__device__ __forceinline__ double div_exact(double a, double b, double y) {
	const double q0 = dmul(a, y);
	const double r = __fma_rn(-b, q0, a);
	return __fma_rn(r, y, q0);
}

// SMALLN: the caller guarantees n <= SSC_SMALLN_MAX and passes the tables (no fallback code then)
template <bool SMALLN>
__device__ __forceinline__ double sd_from_sums(unsigned long long S, unsigned long long Q, int n,
		const SmallN *t = nullptr) {
	const unsigned long long D = (unsigned long long) n * Q - S * S;
	double var, nn1;
	if constexpr (SMALLN) {
		// D < n^2 * 2^32 <= 2^52 (n <= 1024): exact through the 2^52 mantissa trick, no conversion unit
		const double d = dsub(__hiloint2double(0x43300000 | (int) (D >> 32), (int) (unsigned int) D), 4503599627370496.0);
		var = div_exact(d, (double) (n * n), t->inv_n2[n]);
		nn1 = t->unbias[n];
	} else {
		var = ddiv(__ull2double_rn(D), (double) ((long long) n * n));
		nn1 = ddiv((double) n, (double) (n - 1));
	}
	return dsqrt(dmul(var, nn1));
}

// stacking.c:1794: round_to_WORD(sum / (double) N), the sum being an exact integer
template <bool SMALLN>
__device__ __forceinline__ unsigned int mean_to_word(unsigned long long S, int n, const SmallN *t = nullptr) {
	const double s = (double) (long long) S;
	if constexpr (SMALLN) return round_to_word(div_exact(s, (double) n, t->inv_n[n]));
	else return round_to_word(ddiv(s, (double) n));
}

// sigma_clipping, stacking.c:1148-1161: -1 low, +1 high, 0 keep
__device__ __forceinline__ int clip_sigma(unsigned int p, double slow, double shigh, double sigma,
		double median) {
	double x = (double) (int) p;
	if (dsub(median, x) > dmul(slow, sigma)) return -1;
	if (dsub(x, median) > dmul(shigh, sigma)) return 1;
	return 0;
}

// percentile_clipping, stacking.c:1130-1143
__device__ __forceinline__ int clip_percentile(unsigned int p, double plow, double phigh, double median) {
	double x = (double) (int) p;
	if (ddiv(dsub(median, x), median) > plow) return -1;
	if (ddiv(dsub(x, median), median) > phigh) return 1;
	return 0;
}

// a / b correctly rounded for a divisor that is used many times: y = RN(1 / b) is computed once
// (one real division), then Markstein's sequence as in div_exact().  The theorem excludes divisors
// whose significand is all ones, and the intermediate products must stay normal: anything else
// (b zero, subnormal, huge, NaN ...) keeps the plain division.
struct Reciprocal {
	double b, y;
	bool ok;
	__device__ __forceinline__ explicit Reciprocal(double divisor) : b(divisor) {
		const unsigned int hi = (unsigned int) __double2hiint(divisor), lo = (unsigned int) __double2loint(divisor);
		const unsigned int e = (hi >> 20) & 0x7FFu;
		ok = e > 1023u - 200u && e < 1023u + 200u && !((hi & 0xFFFFFu) == 0xFFFFFu && lo == 0xFFFFFFFFu);
		y = ok ? ddiv(1.0, divisor) : 0.0;
	}
	// |a| must be 0 or within 2^+-200 too (the callers divide pixel-scale residuals)
	__device__ __forceinline__ double divide(double a) const {
		if (!ok) return ddiv(a, b);
		const double q0 = dmul(a, y);
		const double r = __fma_rn(-b, q0, a);
		return __fma_rn(r, y, q0);
	}
};

// line_clipping, stacking.c:1170-1183
__device__ __forceinline__ int clip_line(const Reciprocal &inv_sigma, unsigned int p, double slow, double shigh, int i,
		double a, double b) {
	double x = (double) (int) p, xi = (double) i;
	if (inv_sigma.divide(dsub(dadd(dmul(a, xi), b), x)) > slow) return -1;
	if (inv_sigma.divide(dsub(dsub(x, dmul(a, xi)), b)) > shigh) return 1;
	return 0;
}

__device__ __forceinline__ int clip_line(unsigned int p, double slow, double shigh, double sigma, int i,
		double a, double b) {
	double x = (double) (int) p, xi = (double) i;
	if (ddiv(dsub(dadd(dmul(a, xi), b), x), sigma) > slow) return -1;
	if (ddiv(dsub(dsub(x, dmul(a, xi)), b), sigma) > shigh) return 1;
	return 0;
}

}  // namespace ssc
