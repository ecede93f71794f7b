// ssc_network.cuh -- pieces shared by the register-network kernels: the packed u16x2 sorting
// network and the exact FP64 normalization of a packed sample pair.
#pragma once
#include "ssc_device.cuh"
#include "ssc_sortnet.h"

namespace ssc {

__device__ __forceinline__ double word_to_double(unsigned int p) {
	return dsub(__hiloint2double(0x43300000, (int) p), 4503599627370496.0);	// (2^52 + p) - 2^52, exact
}

// round_to_WORD (utils.c:68-74) for |y| < 2^31: floor(y + 0.5) read from the low word of
// (y + 0.5) +rz 1.5*2^52, then max(min(., 65535), 0).  Equals the reference for every such y:
// y <= 0 -> floor <= 0 -> 0;  0 < y <= 65535 -> trunc(y + 0.5);  y > 65535 -> >= 65535 -> 65535.
__device__ __forceinline__ unsigned int round_to_word_fast(double y) {
	double u = __dadd_rz(dadd(y, 0.5), 6755399441055744.0);
	return (unsigned int) __vimin_s32_relu(__double2loint(u), 65535);
}

template <int NP>
__device__ __forceinline__ void sort_network(unsigned int (&v)[NP]) {
#define SSC_CE(a, b) { unsigned int lo_ = __vminu2(v[a], v[b]); v[b] = __vmaxu2(v[a], v[b]); v[a] = lo_; }
	if constexpr (NP == 16) { SSC_SORTNET_16 }
	else if constexpr (NP == 32) { SSC_SORTNET_32 }
	else { SSC_SORTNET_64 }
#undef SSC_CE
}

// The same network with every comparator half removed that cannot reach the ranks the register sigma
// clipping reads (ssc_stack_fast.cu MODE_SIGMA_REG: 0..2, NP/2-3..NP/2+1, NP-6..NP-1): 888 instead of
// 1086 packed min/max operations for 64 frames.  Only those ranks are valid afterwards.
template <int NP>
__device__ __forceinline__ void sort_network_sigma_ranks(unsigned int (&v)[NP]) {
#define SSC_CE(a, b) { unsigned int lo_ = __vminu2(v[a], v[b]); v[b] = __vmaxu2(v[a], v[b]); v[a] = lo_; }
#define SSC_MIN(a, b) { v[a] = __vminu2(v[a], v[b]); }
#define SSC_MAX(a, b) { v[b] = __vmaxu2(v[a], v[b]); }
	if constexpr (NP == 32) { SSC_SORTNET_32_SIGMA }
	else { SSC_SORTNET_64_SIGMA }
#undef SSC_CE
#undef SSC_MIN
#undef SSC_MAX
}

// One sample pair through normalization (stacking.c:1635-1652), KIND = the reference's enum value.
// ADDITIVE / MULTIPLICATIVE have scale == 1.0 exactly (stacking.c:130-134), and x * 1.0 == x.
template <int KIND>
__device__ __forceinline__ unsigned int normalize_pair_fast(unsigned int w, double off, double mul, double sc,
		unsigned long long &Qa, unsigned long long &Qb) {
	// I2F.F64.U16 reads either half of the packed word directly (conversion unit, off the ALU pipe)
	double da = (double) (unsigned short) (w & 0xFFFFu), db = (double) (unsigned short) (w >> 16);
	if constexpr (KIND == ADDITIVE_SCALING || KIND == MULTIPLICATIVE_SCALING) {
		da = dmul(da, sc);
		db = dmul(db, sc);
	}
	if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) {
		da = dsub(da, off);
		db = dsub(db, off);
	} else {
		da = dmul(da, mul);
		db = dmul(db, mul);
	}
	const unsigned int ra = round_to_word_fast(da), rb = round_to_word_fast(db);
	Qa += (unsigned long long) ra * ra;	// IMAD.WIDE on the values while they are still unpacked
	Qb += (unsigned long long) rb * rb;
	return rb * 65536u + ra;	// one IMAD on the FMA pipe (= ra | rb << 16, both below 2^16): the ALU pipe is the busy one
}

// ---------------------------------------------------------------------------------------------
// The CHEAP form of the normalization: one fused multiply-add per sample instead of a conversion,
// a product and two additions.
//   multiplicative kinds:  u = fma(2^52 + p, A, C),  A = mul (or RN(scale * mul)),  C = 2^52 - 2^52*A
//                          = 2^52 + RN_int(p * A): the low word of u is p * A rounded to the nearest integer
//                          (ties to even), and for A >= 0 its high word is 0x43300000 again;
//   additive kinds:        u = (fma(2^52 + p, scale, -2^52*scale) - offset) + 1.5*2^52
//                          -- the fma IS RN(p * scale), the subtraction the reference's, only the rounding to an
//                          integer differs (nearest-even instead of trunc(. + 0.5)).
// 2^52 + p costs nothing: it is the register pair {p, 0x43300000} -- p straight out of a 16-bit shared-memory
// load (no extraction, no conversion unit), and in the multiplicative kinds the high word is the one the previous
// sample's result left in the odd register of the pair (cheap_sample's hi52 goes in and comes out).
// The reference rounds every operation individually and rounds halves up (stacking.c:1635-1652,
// utils.c:68-74), so the two forms can differ -- for a given frame only at samples p whose product lies within
// ~2^-37 of a half-integer.  They are therefore only ever used after k_verify_cheap_norm
// (ssc_stack_generic.cu) has compared them with normalize_sample() for ALL 65536 sample values of EVERY
// frame of the launch: one differing value anywhere and the launch keeps the exact form.
// Used by the 65..1024-frame kernel (ssc_stack_mid.cu: cfg4 19.2 -> 18.8 ms).  The 64-frame register kernel keeps the
// exact form: its normalization loop is bound by the conversion unit, the FP64 pipe, the FMA pipe (IMAD.WIDE) and the
// shared-memory port all at once (tools/micro/pipes.cu), and relieving two of them moved nothing.
// ---------------------------------------------------------------------------------------------
constexpr double SSC_TWO52 = 4503599627370496.0, SSC_MAGIC = 6755399441055744.0;	// 2^52, 1.5 * 2^52
constexpr unsigned int SSC_HI52 = 0x43300000u;	// high word of the doubles in [2^52, 2^52 + 2^32)

// the two per-frame constants of the cheap form (a, c) from the reference's coefficients
template <int KIND>
__device__ __forceinline__ void cheap_constants(double off, double mul, double sc, double &a, double &c) {
	if constexpr (KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING) {
		a = KIND == MULTIPLICATIVE_SCALING ? dmul(sc, mul) : mul;
		c = __fma_rn(-SSC_TWO52, a, SSC_TWO52);
	} else {
		a = KIND == ADDITIVE_SCALING ? sc : 1.0;
		c = dmul(-SSC_TWO52, a);
	}
}

// One sample p < 65536 through the cheap form (off is only read by the additive kinds).  hi52 must be SSC_HI52 on
// entry; the multiplicative kinds hand the high word of their result back in it -- SSC_HI52 again whenever
// 0 <= p * a < 2^31, which k_verify_cheap_norm checks for every p -- so that a rolled loop carries it from sample
// to sample in a register instead of moving the constant into place once per sample.
template <int KIND>
__device__ __forceinline__ unsigned int cheap_sample(unsigned int p, double a, double c, double off, unsigned int &hi52) {
	const double d = __hiloint2double((int) hi52, (int) p);
	double u = __fma_rn(d, a, c);
	if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) {
		u = dadd(dsub(u, off), SSC_MAGIC);
		return (unsigned int) __vimin_s32_relu(__double2loint(u), 65535);
	} else {
		hi52 = (unsigned int) __double2hiint(u);
		return (unsigned int) __vimin_s32_relu(__double2loint(u), 65535);	// (signed + RELU: the compiler must not learn the range)
	}
}

// what k_verify_cheap_norm compares with normalize_sample(): the value, and whether the high word came back intact
__device__ __forceinline__ unsigned int cheap_sample_dyn(unsigned int p, int mode, double off, double mul, double sc, bool &hi_ok) {
	double a, c;
	unsigned int hi = SSC_HI52, r;
	switch (mode) {
	case ADDITIVE: cheap_constants<ADDITIVE>(off, mul, sc, a, c); r = cheap_sample<ADDITIVE>(p, a, c, off, hi); break;
	case ADDITIVE_SCALING: cheap_constants<ADDITIVE_SCALING>(off, mul, sc, a, c); r = cheap_sample<ADDITIVE_SCALING>(p, a, c, off, hi); break;
	case MULTIPLICATIVE: cheap_constants<MULTIPLICATIVE>(off, mul, sc, a, c); r = cheap_sample<MULTIPLICATIVE>(p, a, c, off, hi); break;
	case MULTIPLICATIVE_SCALING: cheap_constants<MULTIPLICATIVE_SCALING>(off, mul, sc, a, c); r = cheap_sample<MULTIPLICATIVE_SCALING>(p, a, c, off, hi); break;
	default: r = p; break;
	}
	hi_ok = hi == SSC_HI52;
	return r;
}

// the pair of samples a thread owns: pa, pb < 65536 as they come out of two 16-bit loads
template <int KIND>
__device__ __forceinline__ unsigned int cheap_pair(unsigned int pa, unsigned int pb, double a, double c, double off,
		unsigned long long &Qa, unsigned long long &Qb, unsigned int &hi_a, unsigned int &hi_b) {
	const unsigned int ra = cheap_sample<KIND>(pa, a, c, off, hi_a), rb = cheap_sample<KIND>(pb, a, c, off, hi_b);
	Qa += (unsigned long long) ra * ra;
	Qb += (unsigned long long) rb * rb;
	return rb * 65536u + ra;	// one IMAD on the FMA pipe (= ra | rb << 16)
}

// true when the frame's coefficients keep every intermediate of the 2^52 forms (exact and cheap) below 2^31
__device__ __forceinline__ bool norm_domain_ok(int mode, double off, double mul, double sc) {
	const bool additive = mode == ADDITIVE || mode == ADDITIVE_SCALING;
	const double bound = additive ? dadd(dmul(fabs(sc), 65535.0), fabs(off)) : dmul(dmul(fabs(sc), 65535.0), fabs(mul));
	return bound < 2147483000.0;	// (false for NaN / inf)
}

// 16-bit shared-memory load, zero-extended (never merged into a 32-bit load + extraction by the compiler)
__device__ __forceinline__ unsigned int lds_u16(const void *ptr) {
	unsigned short v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"((unsigned int) __cvta_generic_to_shared(ptr)) : "memory");
	return v;
}

}  // namespace ssc
