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

}  // namespace ssc
