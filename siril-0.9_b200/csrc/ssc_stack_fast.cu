// ssc_stack_fast.cu -- K2/K3 for stacks of up to 64 frames without registration
// shifts: the shape of the headline workloads (32-frame median, 64-frame
// sigma-clipped mean).
//
// One thread owns TWO horizontally adjacent pixels.  Their samples travel
// packed as u16x2 in one 32-bit register per frame -- which is exactly how two
// adjacent WORDs sit in memory, so a frame costs one coalesced 4-byte access per
// thread (128 B per warp) and no repacking.  The NP (16/32/64) registers are
// sorted by a Batcher odd-even merge network of VIMNMX.U16x2 min/max pairs:
// every compare-exchange orders both pixels at once, every index is a
// compile-time constant, nothing leaves the register file.
//
// Instruction-cache economy decides the rest (tools/micro/ifetch.cu):
// an SM serves ~32 KB of straight-line code at full rate, and a 64-frame tile
// with an unrolled gather and normalization is 48 KB.  So 64-frame stacks are
// STAGED: one cp.async per frame into the thread's shared-memory column, a
// rolled loop that normalizes the column in place and takes the exact sums,
// NP static loads into the registers of the network.  Stacks of <= 32 frames
// keep the fully unrolled form (their code is small).
//
// Sigma clipping on full stacks of 32/64 frames (MODE_SIGMA_REG) never leaves
// the registers for most pixels: a conservative single-precision test on the
// extremes proves that nothing is rejected (clip_cannot_fire); the others are
// packed into 48-byte items on a per-warp shared-memory queue and evaluated
// exactly in FP64, one pixel per lane with full warps, one clipping pass per
// round.  The other modes park the sorted columns in shared memory, where the
// iterative, data-dependent rejection code (ssc_column.cuh) indexes them
// dynamically through a per-block work queue.
//
// Normalization (stacking.c:1635-1652) keeps the reference's individually
// rounded FP64 operations but avoids the slow double->int conversion unit:
// round_to_WORD through an add in round-toward-zero mode onto 1.5*2^52 (floor
// in the low word) and one VIMNMX.RELU clamp.  Exact; frames whose coefficients
// would leave the trick's domain (|value| >= 2^31) take the conversion-based
// form instead.
#include <cstdlib>
#include "ssc_column.cuh"
#include "ssc_kernels.h"
#include "ssc_network.cuh"
#include "ssc_leftover.cuh"

namespace ssc {


template <int NP, bool FULL, int KIND>
__device__ __forceinline__ void normalize_all(unsigned int (&v)[NP], int N, const double *s_coef,
		unsigned long long &Qa, unsigned long long &Qb) {
#pragma unroll
	for (int f = 0; f < NP; f++) {
		if (FULL || f < N) {
			double off = 0.0, mul = 1.0, sc = 1.0;
			if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) off = s_coef[f];
			if constexpr (KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING) mul = s_coef[64 + f];
			if constexpr (KIND == ADDITIVE_SCALING || KIND == MULTIPLICATIVE_SCALING) sc = s_coef[128 + f];
			v[f] = normalize_pair_fast<KIND>(v[f], off, mul, sc, Qa, Qb);
		}
	}
}

enum { MODE_MEDIAN = 0, MODE_NOREJ = 1, MODE_SIGMA_REG = 2, MODE_QUEUE = 3 };

// MODE_SIGMA_REG stages the band through shared memory so that everything before the sorting
// network is ROLLED code: the straight-line part of a tile then stays under the 32 KB of
// instructions an SM serves at full rate (tools/micro/ifetch.cu: beyond that, issue drops to
// ~0.3-0.6 warp-instructions per clock and scheduler whatever the pipes could do).
// One thread's staging column: stage[f * FAST_THREADS + tid], conflict-free.
template <int NP, int KIND, bool FULL>
__device__ __forceinline__ void normalize_staged(unsigned int *stage_t, int stride, const double *s_coef, int N,
		unsigned int &Sa, unsigned int &Sb, unsigned long long &Qa, unsigned long long &Qb) {
	constexpr int G = 8;	// frames per round
	static_assert(NP % G == 0, "groups of G frames");
#pragma unroll 1
	for (int f0 = 0; f0 < (FULL ? NP : N); f0 += G) {
		// all loads of a group first: the coefficient loads are off the critical path of the FP64 chain
		unsigned int w[G];
		double off[G], mul[G], sc[G];
#pragma unroll
		for (int j = 0; j < G; j++) {
			w[j] = stage_t[(f0 + j) * stride];
			off[j] = 0.0; mul[j] = 1.0; sc[j] = 1.0;
			if constexpr (KIND == ADDITIVE || KIND == ADDITIVE_SCALING) off[j] = s_coef[f0 + j];
			if constexpr (KIND == MULTIPLICATIVE || KIND == MULTIPLICATIVE_SCALING) mul[j] = s_coef[64 + f0 + j];
			if constexpr (KIND == ADDITIVE_SCALING || KIND == MULTIPLICATIVE_SCALING) sc[j] = s_coef[128 + f0 + j];
		}
#pragma unroll
		for (int j = 0; j < G; j++) {
			if (!FULL && f0 + j >= N) break;	// (the words behind the N samples hold the 0xFFFF sentinels)
			if constexpr (KIND != NO_NORM) {
				w[j] = normalize_pair_fast<KIND>(w[j], off[j], mul[j], sc[j], Qa, Qb);
				stage_t[(f0 + j) * stride] = w[j];
			} else {
				const unsigned int pa = w[j] & 0xFFFFu, pb = w[j] >> 16;
				Qa += (unsigned long long) pa * pa;
				Qb += (unsigned long long) pb * pb;
			}
			Sa = __dp2a_lo(w[j], 0x0001u, Sa);		// IDP.2A: acc + lo16 * 1 + hi16 * 0
			Sb = __dp2a_lo(w[j], 0x0100u, Sb);
		}
	}
}

// Coefficients outside the domain of the fast form (|value| >= 2^31): the conversion-based one.
__device__ __noinline__ void normalize_staged_slow(unsigned int *stage_t, int stride, const double *s_coef, int N, int mode,
		unsigned int &Sa, unsigned int &Sb, unsigned long long &Qa, unsigned long long &Qb) {
	for (int f = 0; f < N; f++) {
		const unsigned int w = stage_t[f * stride];
		const unsigned int pa = normalize_sample(w & 0xFFFFu, mode, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
		const unsigned int pb = normalize_sample(w >> 16, mode, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
		stage_t[f * stride] = pa | (pb << 16);
		Sa += pa; Sb += pb;
		Qa += (unsigned long long) pa * pa;
		Qb += (unsigned long long) pb * pb;
	}
}

// True when the first sigma-clipping pass (stacking.c:1674-1690) provably rejects nothing in a
// full stack of NP samples.  The exact test is  d_lo > floor(X_lo)  /  d_hi > floor(X_hi)  with
// X = 2 * sig * sigma in FP64 (see the queue pass below) and d_lo = M2 - 2 min, d_hi = 2 max - M2
// (M2 = twice the median; if the extremes stay, everything stays).  Here X is estimated in FP32
// from the exact integer D = n Q - S^2 (< 2^44): conversion, two products, an approximate square
// root and the product with the parameter are each good to ~1e-7 relative, so t * (1 - 1e-4) is
// below X, and  d < t (1 - 1e-4) - 1  implies  d < X - 1 < floor(X).  d is an integer below 2^24,
// exact in FP32.  Anything not provably clear is evaluated exactly.
template <int NP>
__device__ __forceinline__ bool clip_cannot_fire(unsigned long long S, unsigned long long Q, int d_lo, int d_hi,
		float sig_low_f, float sig_high_f) {
	const unsigned long long D = (unsigned long long) NP * Q - S * S;
	float sigma;
	asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sigma) : "f"(__ull2float_rn(D) * (1.0f / ((float) NP * (float) (NP - 1)))));
	const float t_lo = 2.0f * sig_low_f * sigma * 0.9999f - 1.0f, t_hi = 2.0f * sig_high_f * sigma * 0.9999f - 1.0f;
	return (d_lo == 0 || (float) d_lo < t_lo) && (d_hi == 0 || (float) d_hi < t_hi);
}

constexpr int FAST_QUEUE_WORDS = 12, FAST_QUEUE_CAP = 96;	// per warp: < 32 waiting + up to 64 new per tile

// FULL: n_frames == NP (no sentinel padding, no per-frame bound checks in the unrolled code)
// MODE: which tail follows the network (one per instantiation: the kernel is instruction-fetch
//       sensitive, dead alternatives are not carried along)
constexpr int FAST_THREADS = 128;	// 4 blocks per SM (measured: 256/512-thread blocks, tile barriers and warp ping-pong are no faster)

// ROWTILE (MODE_SIGMA_REG only): a warp's 32 pixel pairs are consecutive in ONE row and rows are
// 16-byte aligned, so the gather is sixteen 16-byte copies per thread and tile (each lane fetches four
// pairs of a frame for its neighbours) instead of 64 4-byte copies with their address arithmetic
template <int NP, bool FULL, int MODE, bool ROWTILE = false>
__global__ void __launch_bounds__(FAST_THREADS, 4)
k_stack_fast(StackParams p, LeftoverList list, unsigned int n_tiles) {
	extern __shared__ __align__(16) unsigned int cols[];	// [NP][FAST_THREADS] packed columns
	__shared__ double s_coef[3 * 64];
	__shared__ double s_unbias[65], s_inv_n[65], s_inv_n2[65];	// small-n tables (ssc_device.cuh)
	__shared__ int s_fast_norm;
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	__shared__ unsigned int s_qcount;			// work queue of pixels whose first pass rejects something
	__shared__ unsigned short s_qitem[2 * FAST_THREADS], s_res[2 * FAST_THREADS];
	__shared__ unsigned char s_layer[FAST_THREADS];
	__shared__ unsigned int s_S[2 * FAST_THREADS];
	__shared__ unsigned long long s_Q[2 * FAST_THREADS];


	// STAGED: the samples travel through a shared-memory staging column and everything before the
	// network is rolled code (64-frame stacks: their unrolled form overflows the instruction cache);
	// PREFETCH: the next tile's copies are issued under the network (the staging buffer is free then)
	constexpr bool STAGED = MODE == MODE_SIGMA_REG || NP == 64;
	constexpr bool PREFETCH = MODE == MODE_SIGMA_REG;

	const ssc_band &b = p.band;
	const int W = b.width, N = FULL ? NP : b.n_frames, tid = threadIdx.x, lane = tid & 31;
	const int W2 = W >> 1;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;
	// MODE_SIGMA_REG: the dynamic shared memory holds the per-warp queues of pixels whose clipping
	// has to be evaluated exactly, [warp][FAST_QUEUE_WORDS][FAST_QUEUE_CAP]
	// behind the staging buffer [NP][FAST_THREADS] (see normalize_staged)
	unsigned int (*queue)[FAST_QUEUE_CAP] = reinterpret_cast<unsigned int (*)[FAST_QUEUE_CAP]>(cols + NP * FAST_THREADS)
			+ (tid >> 5) * FAST_QUEUE_WORDS;
	int qcount = 0;		// items on this warp's queue (warp-uniform)
	unsigned int lane_rej[6] = {0, 0, 0, 0, 0, 0};	// this lane's share of irej[3][2] (MODE_SIGMA_REG)
	const float sig_low_f = (float) p.sig_low, sig_high_f = (float) p.sig_high;

	if (tid < SSC_COUNTERS) blk_counters[tid] = 0;
	if (tid == 0) s_fast_norm = 1;
	if (tid >= 1 && tid <= 64) {
		s_unbias[tid] = tid > 1 ? ddiv((double) tid, (double) (tid - 1)) : 0.0;
		s_inv_n[tid] = ddiv(1.0, (double) tid);
		s_inv_n2[tid] = ddiv(1.0, (double) (tid * tid));
	}
	const SmallN tables{s_unbias, s_inv_n, s_inv_n2};
	__syncthreads();
	if (norm && tid < N) {
		const double off = p.coeff[tid], mul = p.coeff[N + tid], sc = p.coeff[2 * N + tid];
		s_coef[tid] = off;
		s_coef[64 + tid] = mul;
		s_coef[128 + tid] = sc;
		// domain of the 2^52 tricks: every intermediate stays below 2^31 in magnitude
		const bool additive = p.normalize == ADDITIVE || p.normalize == ADDITIVE_SCALING;
		const double bound = additive ? dadd(dmul(fabs(sc), 65535.0), fabs(off)) : dmul(dmul(fabs(sc), 65535.0), fabs(mul));
		if (!(bound < 2147483000.0)) s_fast_norm = 0;	// also catches NaN / inf
	}
	__syncthreads();

	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * W2;
	// MODE_SIGMA_REG: request a tile's samples (see the gather below); pf_* describe the requested tile
	int pf_layer = 0;
	unsigned int pf_out_idx = 0;
	bool pf_active = false;
	const bool punt_all = MODE == MODE_SIGMA_REG && norm && !s_fast_norm;
	const int warp = tid >> 5;
	const unsigned int xblocks = ((unsigned int) W2 + 31u) >> 5;	// ROWTILE: warp tiles per row
	// the thread's staging column: element f at stage_t[f * stage_stride]
	unsigned int *const stage_t = ROWTILE ? cols + warp * NP * 32 + lane : cols + tid;
	constexpr int stage_stride = ROWTILE ? 32 : FAST_THREADS;
	// where tile t puts this thread: layer, result row (relative), pixel pair in the row
	auto locate = [&](unsigned long long t, int &lay, int &rr, int &x2, int &x_warp) -> bool {
		// (fast_path_eligible guarantees 32-bit pixel counts and frame strides: no 64-bit division)
		if constexpr (ROWTILE) {
			const unsigned int wt = (unsigned int) t;	// (ROWTILE: tiles are warp tiles, handed out one by one)
			const unsigned int n_wt = (unsigned int) b.layers * (unsigned int) b.out_rows * xblocks;
			const unsigned int wc = wt < n_wt ? wt : n_wt - 1;
			const unsigned int row = wc / xblocks, xb = wc - row * xblocks;
			lay = (int) (row / (unsigned int) b.out_rows);
			rr = (int) (row - (unsigned int) lay * (unsigned int) b.out_rows);
			x_warp = (int) (xb * 32);		// the warp's first pair
			x2 = x_warp + lane;
			const bool ok = wt < n_wt && x2 < W2;
			if (x2 >= W2) x2 = W2 - 1;
			return ok;
		} else {
			unsigned long long gg = t * FAST_THREADS + tid;
			const bool ok = gg < total;
			if (!ok) gg = total - 1;
			const unsigned int g32 = (unsigned int) gg, row = g32 / (unsigned int) W2;
			x2 = (int) (g32 - row * (unsigned int) W2);
			lay = (int) (row / (unsigned int) b.out_rows);
			rr = (int) (row - (unsigned int) lay * (unsigned int) b.out_rows);
			x_warp = 0;
			return ok;
		}
	};
	auto issue_tile = [&](unsigned long long t) {
		int lay, rr, x2, x_warp;
		pf_active = locate(t, lay, rr, x2, x_warp);
		const int R = b.out_row0 + rr;				// result storage row (stacking.c:731,1600)
		const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
		pf_layer = lay;
		pf_out_idx = (unsigned int) ((size_t) lay * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W + 2 * x2);
		const unsigned int fstride32 = (unsigned int) (b.frame_stride * sizeof(WORD));
		const size_t row_word = ((size_t) lay * b.layer_stride + (size_t) (srow - b.buf_row0) * W) >> 1;	// first pair of the row
		if constexpr (ROWTILE) {
			// lane = 8 * s + c fetches pairs [4c, 4c + 4) of the warp's 32 for frames s, s + 4, s + 8, ...
			const int c = lane & 7, sub = lane >> 3;
			const int xw = x_warp + 4 * c;			// first pair of my chunk
			const bool chunk_ok = xw < W2;			// W2 % 4 == 0: a chunk is inside the row or outside
			const char *src0 = reinterpret_cast<const char *>(reinterpret_cast<const unsigned int *>(b.frames) + row_word + xw)
					+ (unsigned long long) ((unsigned int) sub) * fstride32;
			const unsigned int sbase = (unsigned int) __cvta_generic_to_shared(cols + warp * NP * 32 + sub * 32 + 4 * c);
			if (chunk_ok) {
#pragma unroll 4
				for (int i = 0; i < NP / 4; i++) {
					const char *src = src0 + (unsigned long long) ((unsigned int) (4 * i)) * fstride32;
					asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sbase + i * 4 * 32 * (int) sizeof(unsigned int)), "l"(src) : "memory");
				}
			}
		} else {
			const char *src0 = reinterpret_cast<const char *>(reinterpret_cast<const unsigned int *>(b.frames) + row_word + x2);
			const unsigned int sbase = (unsigned int) __cvta_generic_to_shared(cols + tid);
#pragma unroll 4
			for (int k = 0; k < NP; k++) {
				if (FULL || k < N) {
					const char *src = src0 + (unsigned long long) ((unsigned int) k) * fstride32;	// one IMAD.WIDE
					asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sbase + k * FAST_THREADS * (int) sizeof(unsigned int)), "l"(src) : "memory");
				} else {
					cols[k * FAST_THREADS + tid] = 0xFFFFFFFFu;	// sentinels sort behind the N samples
				}
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	// ROWTILE: every warp draws its next warp tile from the launch's work counter, so a launch ends with at most
	// one tile per warp outstanding whatever its size (a band of 1/8 of the image left a partial wave of
	// statically assigned tiles: 0.60 ms for the work of 0.54)
	auto next_after = [&](unsigned long long t) -> unsigned long long {
		if constexpr (ROWTILE) {
			unsigned int n = 0;
			if (lane == 0) n = atomicAdd(list.next_tile, 1u);
			return __shfl_sync(0xffffffffu, n, 0);
		} else {
			return t + gridDim.x;
		}
	};
	unsigned long long first_tile = ROWTILE ? next_after(0) : blockIdx.x;
	if (PREFETCH && !punt_all && first_tile < n_tiles) issue_tile(first_tile);
	// Blocks loop over tiles (the launcher bounds the grid): the tables, the coefficient copy and
	// their barriers above are paid once per block instead of once per tile.
	// (MODE_SIGMA_REG runs one more round after its last tile, to drain the warp queues)
	unsigned long long upcoming = 0;
	for (unsigned long long tile = first_tile; ; tile = upcoming) {
	const bool more = tile < n_tiles;
	if (MODE != MODE_SIGMA_REG && !more) break;
	upcoming = more ? next_after(tile) : tile;
	if (more) {
	unsigned long long g = tile * FAST_THREADS + tid;
	bool active = g < total;
	if constexpr (MODE == MODE_SIGMA_REG) {
		if (punt_all) {	// coefficients outside the fast normalization's domain (block-uniform):
			int lay, rr, x2, x_warp;
			const bool ok = locate(tile, lay, rr, x2, x_warp);
			const unsigned int pix = (unsigned int) (((size_t) lay * b.out_rows + rr) * W + 2 * x2);
			leftover_push(list, ok, pix, lane);		// every pixel goes to the general kernel
			leftover_push(list, ok, pix + 1, lane);
			continue;
		}
	}
	int layer;
	unsigned int out_idx;
	unsigned int v[NP];
	unsigned long long Qa = 0, Qb = 0;	// exact sums of squares of the (normalized) samples
	unsigned int Sa = 0, Sb = 0;		// exact sums (fit 32 bits)
	bool have_q = false;
	const char *ptr = nullptr;
	size_t fstride_bytes = 0;
	if constexpr (STAGED) {
		// ---- asynchronous 4-byte copies, one per frame, straight into the thread's staging column
		// (issue_tile; with PREFETCH they were issued a tile ago).  Normalization and the sums run in
		// a rolled loop over the column, the network gets it through NP static loads -- and with
		// PREFETCH the NEXT tile's copies are issued as soon as the column is in registers, to land
		// under the network and the rejection.
		if constexpr (!PREFETCH) issue_tile(tile);
		layer = pf_layer;
		out_idx = pf_out_idx;
		active = pf_active;
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		if constexpr (ROWTILE) __syncwarp();	// the lanes fetched each other's samples
		if (norm && !s_fast_norm) {
			normalize_staged_slow(stage_t, stage_stride, s_coef, N, p.normalize, Sa, Sb, Qa, Qb);
		} else {
			switch (norm ? p.normalize : NO_NORM) {
			case NO_NORM: normalize_staged<NP, NO_NORM, FULL>(stage_t, stage_stride, s_coef, N, Sa, Sb, Qa, Qb); break;
			case ADDITIVE: normalize_staged<NP, ADDITIVE, FULL>(stage_t, stage_stride, s_coef, N, Sa, Sb, Qa, Qb); break;
			case MULTIPLICATIVE: normalize_staged<NP, MULTIPLICATIVE, FULL>(stage_t, stage_stride, s_coef, N, Sa, Sb, Qa, Qb); break;
			case ADDITIVE_SCALING: normalize_staged<NP, ADDITIVE_SCALING, FULL>(stage_t, stage_stride, s_coef, N, Sa, Sb, Qa, Qb); break;
			default: normalize_staged<NP, MULTIPLICATIVE_SCALING, FULL>(stage_t, stage_stride, s_coef, N, Sa, Sb, Qa, Qb); break;
			}
		}
#pragma unroll
		for (int k = 0; k < NP; k++) v[k] = stage_t[k * stage_stride];
		have_q = true;
		if constexpr (ROWTILE) __syncwarp();	// everybody has its column: the buffer may be refilled
		if (PREFETCH && upcoming < n_tiles) issue_tile(upcoming);
	} else {
	if (!active) g = total - 1;
	const int x2 = (int) (g % W2);
	const int rr = (int) ((g / W2) % b.out_rows);
	layer = (int) (g / ((unsigned long long) W2 * b.out_rows));
	const int R = b.out_row0 + rr;				// result storage row (stacking.c:731,1600)
	const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
	// per-thread offset inside a frame (32 bits: a frame holds < 2^32 pixel pairs) on top of a
	// block-uniform per-frame base, so the 64 address computations stay off the ALU pipe
	const unsigned int toff = (unsigned int) ((((size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W) >> 1) + x2);
	out_idx = (unsigned int) ((size_t) layer * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W + 2 * x2);
	ptr = reinterpret_cast<const char *>(reinterpret_cast<const unsigned int *>(b.frames) + toff);
	fstride_bytes = (b.frame_stride >> 1) * sizeof(unsigned int);
	// ---- gather: one packed load per frame, all issued before first use ----
#pragma unroll
	for (int f = 0; f < NP; f++) {
		v[f] = (FULL || f < N) ? __ldg(reinterpret_cast<const unsigned int *>(ptr)) : 0xFFFFFFFFu;
		ptr += fstride_bytes;
	}

	// ---- normalize (stacking.c:1635-1652; median path :750-764) ----
	if (norm) {
		if (s_fast_norm) {
			have_q = true;
			switch (p.normalize) {
			case ADDITIVE: normalize_all<NP, FULL, ADDITIVE>(v, N, s_coef, Qa, Qb); break;
			case MULTIPLICATIVE: normalize_all<NP, FULL, MULTIPLICATIVE>(v, N, s_coef, Qa, Qb); break;
			case ADDITIVE_SCALING: normalize_all<NP, FULL, ADDITIVE_SCALING>(v, N, s_coef, Qa, Qb); break;
			default: normalize_all<NP, FULL, MULTIPLICATIVE_SCALING>(v, N, s_coef, Qa, Qb); break;
			}
		} else {
			// coefficients outside the tricks' domain: conversion-based form, rolled loop over the
			// shared-memory copy (rare: keep it out of the straight-line code)
#pragma unroll
			for (int f = 0; f < NP; f++) cols[f * FAST_THREADS + tid] = v[f];
			for (int f = 0; f < N; f++) {
				unsigned int w = cols[f * FAST_THREADS + tid];
				unsigned int pa = normalize_sample(w & 0xFFFFu, p.normalize, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
				unsigned int pb = normalize_sample(w >> 16, p.normalize, s_coef[f], s_coef[64 + f], s_coef[128 + f]);
				cols[f * FAST_THREADS + tid] = pa | (pb << 16);
			}
#pragma unroll
			for (int f = 0; f < NP; f++) v[f] = cols[f * FAST_THREADS + tid];
		}
	}
	}

	unsigned int res[2];
	WORD *dst = p.out + out_idx;

	if constexpr (MODE == MODE_MEDIAN) {
		sort_network<NP>(v);			// sentinels (0xFFFF) sort behind the N samples
		unsigned int lo, hi;
		if (FULL) {				// stacking.c:765-767: truncating median, straight from registers
			lo = v[NP / 2 - 1];
			hi = v[NP / 2];
		} else {
#pragma unroll
			for (int f = 0; f < NP; f++) cols[f * FAST_THREADS + tid] = v[f];
			lo = cols[((N - 1) >> 1) * FAST_THREADS + tid];
			hi = cols[(N >> 1) * FAST_THREADS + tid];
		}
		res[0] = ((lo & 0xFFFFu) + (hi & 0xFFFFu)) >> 1;
		res[1] = ((lo >> 16) + (hi >> 16)) >> 1;
		if (active) *reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
		continue;
	} else {
		// exact integer sums of the (normalized) samples; the sum fits 32 bits and comes straight
		// from the packed words (IDP.2A: acc + lo16 * 1 + hi16 * 0), the sum of squares needs 64
		if constexpr (!STAGED) {		// (STAGED: summed in the rolled loop)
#pragma unroll
			for (int f = 0; f < NP; f++) {
				if (FULL || f < N) {
					Sa = __dp2a_lo(v[f], 0x0001u, Sa);
					Sb = __dp2a_lo(v[f], 0x0100u, Sb);
				}
			}
		}
		if (!have_q && MODE != MODE_NOREJ) {
#pragma unroll
			for (int f = 0; f < NP; f++) {
				if (FULL || f < N) {
					const unsigned int pa = v[f] & 0xFFFFu, pb = v[f] >> 16;
					Qa += (unsigned long long) pa * pa;
					Qb += (unsigned long long) pb * pb;
				}
			}
		}
		unsigned long long S[2] = {Sa, Sb}, Q[2] = {Qa, Qb};

		if constexpr (MODE == MODE_NOREJ) {
			// stacking.c:1790-1794 with nothing rejected: no ordering needed at all
			res[0] = mean_to_word<true>(S[0], N, &tables);
			res[1] = mean_to_word<true>(S[1], N, &tables);
			if (active) *reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
			continue;
		} else if constexpr (MODE == MODE_SIGMA_REG) {
			// ---- sigma clipping (stacking.c:1674-1695).  In a sorted stack the low flags are a prefix
			// and the high flags a suffix, the survivors stay [nl, NP - nh), and each pass only looks at
			// the sums, the two middle survivors and the next candidates at both ends: v[0..2] below,
			// v[NP-1..NP-6] above, medians within v[NP/2-3..NP/2+1].
			// Most pixels reject nothing, which a conservative single-precision test on the extreme
			// samples proves without any FP64 work (see clip_cannot_fire).  The others are packed into
			// 48-byte items -- the 14 candidates, the sums, the result address -- on the warp's queue
			// and evaluated exactly, one pixel per lane with FULL warps, one pass per round (an item
			// that rejected something goes back on the queue for its next pass).  Whoever leaves the
			// candidate window (a third low, a sixth high) goes to the leftover list.  With NP >= 32
			// and at most 7 rejections the reference's "N - r <= 4" break (:1684) cannot be reached.
			static_assert(FULL && NP >= 32 && (NP & (NP - 1)) == 0, "needs a full stack of 32 or 64 frames");
			sort_network_sigma_ranks<NP>(v);	// only the ranks read below are in place (the hand-back kernel sorts fully)
			bool need[2];
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const unsigned int lo = h ? v[0] >> 16 : v[0] & 0xFFFFu, hi = h ? v[NP - 1] >> 16 : v[NP - 1] & 0xFFFFu;
				const unsigned int ml = h ? v[NP / 2 - 1] >> 16 : v[NP / 2 - 1] & 0xFFFFu;
				const unsigned int mr = h ? v[NP / 2] >> 16 : v[NP / 2] & 0xFFFFu;
				const int M2 = (int) (ml + mr);
				need[h] = active && !clip_cannot_fire<NP>(S[h], Q[h], M2 - 2 * (int) lo, 2 * (int) hi - M2, sig_low_f, sig_high_f);
				res[h] = (unsigned int) ((S[h] + NP / 2) / NP);	// round_to_WORD(S / NP), NP a power of two
			}
			if (active) {
				if (!need[0] && !need[1]) {
					*reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
				} else {
					if (!need[0]) dst[0] = (WORD) res[0];
					if (!need[1]) dst[1] = (WORD) res[1];
				}
			}
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const unsigned int mask = __ballot_sync(0xffffffffu, need[h]);
				if (mask == 0) continue;
				if (need[h]) {
					const int slot = qcount + __popc(mask & ((1u << lane) - 1));
					const unsigned int sel = h ? 0x7632u : 0x5410u;	// this pixel's halves of two registers
					queue[0][slot] = __byte_perm(v[0], v[1], sel);
					queue[1][slot] = __byte_perm(v[2], v[NP / 2 - 3], sel);
					queue[2][slot] = __byte_perm(v[NP / 2 - 2], v[NP / 2 - 1], sel);
					queue[3][slot] = __byte_perm(v[NP / 2], v[NP / 2 + 1], sel);
					queue[4][slot] = __byte_perm(v[NP - 1], v[NP - 2], sel);
					queue[5][slot] = __byte_perm(v[NP - 3], v[NP - 4], sel);
					queue[6][slot] = __byte_perm(v[NP - 5], v[NP - 6], sel);
					queue[7][slot] = (unsigned int) S[h];
					queue[8][slot] = (unsigned int) Q[h];
					queue[9][slot] = (unsigned int) (Q[h] >> 32);
					queue[10][slot] = out_idx + h;
					queue[11][slot] = (unsigned int) layer << 8;	// nl | nh << 4 | layer << 8
				}
				qcount += __popc(mask);
			}
		} else {	// MODE_QUEUE: every other rejection / stack size, through the block's work queue
			sort_network<NP>(v);
#pragma unroll
			for (int f = 0; f < NP; f++)
				if (FULL || f < N) cols[f * FAST_THREADS + tid] = v[f];
			WORD *cols16 = reinterpret_cast<WORD *>(cols);

			// ---- phase 1: can the first pass flag anything at all?  Flags are monotone in the
			// sample value (low flags a prefix, high flags a suffix of the sorted stack), so when
			// neither the smallest nor the largest sample is flagged nothing is, the reference's
			// do-while ends after one pass (n == 0) and the result is the plain mean.
			// (with N <= 4 the reference's "N - r <= 4" break is taken at the very first sample:
			// that pixel must go through the full pass so that the event is counted)
			const bool quick = (p.rejection == SIGMA && N > 4) || p.rejection == SIGMEDIAN || p.rejection == PERCENTILE;
			const unsigned int lo_w = v[0];
			const unsigned int hi_w = FULL ? v[NP - 1] : cols[(N - 1) * FAST_THREADS + tid];
			const unsigned int ml_w = FULL ? v[(NP - 1) / 2] : cols[((N - 1) >> 1) * FAST_THREADS + tid];
			const unsigned int mr_w = FULL ? v[NP / 2] : cols[(N >> 1) * FAST_THREADS + tid];
			bool pending[2];
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const unsigned int lo = h ? lo_w >> 16 : lo_w & 0xFFFFu, hi = h ? hi_w >> 16 : hi_w & 0xFFFFu;
				const unsigned int ml = h ? ml_w >> 16 : ml_w & 0xFFFFu, mr = h ? mr_w >> 16 : mr_w & 0xFFFFu;
				const double median = (N & 1) ? (double) (int) ml : dmul((double) (int) (ml + mr), 0.5);
				bool flagged = true;
				if (quick) {
					if (p.rejection == PERCENTILE) {
						flagged = clip_percentile(lo, p.sig_low, p.sig_high, median) != 0
								|| clip_percentile(hi, p.sig_low, p.sig_high, median) != 0;
					} else {
						const double sigma = sd_from_sums<true>(S[h], Q[h], N, &tables);
						flagged = clip_sigma(lo, p.sig_low, p.sig_high, sigma, median) != 0
								|| clip_sigma(hi, p.sig_low, p.sig_high, sigma, median) != 0;
					}
				}
				pending[h] = flagged && active;	// padding threads of the last block own no pixel
				res[h] = mean_to_word<true>(S[h], N, &tables);
			}
			if (tid == 0) s_qcount = 0;
			s_layer[tid] = (unsigned char) layer;
			__syncthreads();
#pragma unroll
			for (int h = 0; h < 2; h++) {
				if (pending[h]) {
					const unsigned int slot = atomicAdd(&s_qcount, 1u);
					s_qitem[slot] = (unsigned short) (2 * tid + h);
					s_S[2 * tid + h] = (unsigned int) S[h];
					s_Q[2 * tid + h] = Q[h];
				}
			}
			__syncthreads();

			// ---- phase 2: the pending pixels, densely packed over the block's threads (their sorted
			// columns sit in shared memory, any thread can run them) ----
			const unsigned int n_items = s_qcount;
			for (unsigned int i = tid; i < n_items; i += FAST_THREADS) {
				const unsigned int item = s_qitem[i];
				Column c{cols16 + item, 2 * FAST_THREADS};
				RejectState ist{0, 0, 0, 0};
				const unsigned int r = reject_sorted_and_mean<true>(c, N, p.rejection, p.sig_low, p.sig_high,
						s_S[item], s_Q[item], ist, &tables, &p.fit);
				s_res[item] = (unsigned short) r;
				const int il = s_layer[item >> 1];
				if (ist.rej_lo) atomicAdd(&blk_counters[il * 2 + 0], (unsigned long long) ist.rej_lo);
				if (ist.rej_hi) atomicAdd(&blk_counters[il * 2 + 1], (unsigned long long) ist.rej_hi);
				if (ist.guard) atomicAdd(&blk_counters[6], 1ull);
				if (ist.cap) atomicAdd(&blk_counters[7], 1ull);
			}
			__syncthreads();
			if (pending[0]) res[0] = s_res[2 * tid];
			if (pending[1]) res[1] = s_res[2 * tid + 1];
			if (active) *reinterpret_cast<unsigned int *>(dst) = res[0] | (res[1] << 16);
		}
	}
	if (MODE == MODE_QUEUE) __syncthreads();	// shared-memory columns / queue are reused by the next tile
	}	// if (more)
	if constexpr (MODE == MODE_SIGMA_REG) {
		// ---- exact clipping passes on the warp's queue: full warps while tiles keep coming, whatever
		// is left after the last one.  One pass of one pixel per lane and round.
		const int min_fill = more ? 32 : 1;
		while (qcount >= min_fill) {
			const int cnt = qcount < 32 ? qcount : 32;
			const int slot = qcount - cnt + lane;
			const bool mine = lane < cnt;
			unsigned int w[FAST_QUEUE_WORDS];
#pragma unroll
			for (int k = 0; k < FAST_QUEUE_WORDS; k++) w[k] = mine ? queue[k][slot] : 0u;
			__syncwarp();
			qcount -= cnt;
			const unsigned int low[3] = {w[0] & 0xFFFFu, w[0] >> 16, w[1] & 0xFFFFu};
			const unsigned int mid[5] = {w[1] >> 16, w[2] & 0xFFFFu, w[2] >> 16, w[3] & 0xFFFFu, w[3] >> 16};
			const unsigned int high[6] = {w[4] & 0xFFFFu, w[4] >> 16, w[5] & 0xFFFFu, w[5] >> 16, w[6] & 0xFFFFu, w[6] >> 16};
			unsigned long long S1 = w[7], Q1 = (unsigned long long) w[8] | ((unsigned long long) w[9] << 32);
			int nl = (int) (w[11] & 15u), nh = (int) ((w[11] >> 4) & 15u);
			const int item_layer = (int) (w[11] >> 8);
			const int Nc = NP - nl - nh;
			const double sigma = sd_from_sums<true>(S1, Q1, Nc, &tables);
			const int lhs = nl + ((Nc - 1) >> 1) - (NP / 2 - 3), rhs = nl + (Nc >> 1) - (NP / 2 - 3);
			unsigned int ml = 0, mr = 0;
#pragma unroll
			for (int k = 0; k < 5; k++) {
				if (lhs == k) ml = mid[k];
				if (rhs == k) mr = mid[k];
			}
			// sigma_clipping (:1152,1156) in integers: the median is M2 / 2 with M2 an integer and the
			// sample e an integer, so (median - e) is exact in FP64 and
			//   median - e > tl  <=>  M2 - 2e > 2 tl  <=>  M2 - 2e > floor(2 tl)      (tl >= 0)
			// and likewise on the high side: one float-to-int conversion per pass and side instead
			// of a conversion, a subtraction and a compare per candidate.
			const int M2 = lhs == rhs ? 2 * (int) ml : (int) (ml + mr);
			const int TL = __double2int_rd(fmin(dmul(2.0, dmul(p.sig_low, sigma)), 1073741824.0));
			const int TH = __double2int_rd(fmin(dmul(2.0, dmul(p.sig_high, sigma)), 1073741824.0));
			int n = 0;
#pragma unroll
			for (int k = 0; k < 3; k++) {		// low candidates, in scan order
				const unsigned int e = low[k];
				if (nl == k && M2 - 2 * (int) e > TL) {
					nl++; n++;
					S1 -= e; Q1 -= (unsigned long long) e * e;
				}
			}
#pragma unroll
			for (int k = 0; k < 6; k++) {		// high candidates, from the top
				const unsigned int e = high[k];
				if (nh == k && 2 * (int) e - M2 > TH) {
					nh++; n++;
					S1 -= e; Q1 -= (unsigned long long) e * e;
				}
			}
			const bool punt = mine && (nl == 3 || nh == 6);		// the next candidate is outside the window
			const bool done = mine && !punt && n == 0;		// do { } while (n > 0 && N > 3)
			const bool again = mine && !punt && !done;
			if (done) {
				p.out[w[10]] = (WORD) mean_to_word<true>(S1, NP - nl - nh, &tables);
#pragma unroll
				for (int c = 0; c < 3; c++) {	// per-lane counters (64-bit shared-memory atomics are CAS loops)
					if (item_layer == c) {
						lane_rej[2 * c] += (unsigned int) nl;
						lane_rej[2 * c + 1] += (unsigned int) nh;
					}
				}
			}
			// result index -> index of the pixel in the band (what the general kernel's list holds)
			leftover_push(list, punt, (unsigned int) ((size_t) item_layer * b.out_rows * W
					+ (w[10] - (size_t) item_layer * p.out_layer_stride) - (size_t) (b.out_row0 - p.out_buf_row0) * W), lane);
			const unsigned int mask = __ballot_sync(0xffffffffu, again);
			if (again) {
				const int dst_slot = qcount + __popc(mask & ((1u << lane) - 1));
#pragma unroll
				for (int k = 0; k < 7; k++) queue[k][dst_slot] = w[k];
				queue[7][dst_slot] = (unsigned int) S1;
				queue[8][dst_slot] = (unsigned int) Q1;
				queue[9][dst_slot] = (unsigned int) (Q1 >> 32);
				queue[10][dst_slot] = w[10];
				queue[11][dst_slot] = (unsigned int) nl | ((unsigned int) nh << 4) | ((unsigned int) item_layer << 8);
			}
			qcount += __popc(mask);
			__syncwarp();
		}
		if (!more) break;
	}
	}	// tiles
	if constexpr (MODE == MODE_SIGMA_REG) {
#pragma unroll
		for (int k = 0; k < 6; k++) {
			const unsigned int sum = __reduce_add_sync(0xffffffffu, lane_rej[k]);
			if (lane == 0 && sum) atomicAdd(&blk_counters[k], (unsigned long long) sum);
		}
	}
	__syncthreads();
	if (p.counters && tid < SSC_COUNTERS && blk_counters[tid])
		atomicAdd(&p.counters[tid], blk_counters[tid]);
}

// ---------------------------------------------------------------------------------------------
// The pixels MODE_SIGMA_REG hands back (more rejections than its candidate window holds, or
// coefficients outside the fast normalization's domain): one thread per listed pixel reloads the
// column, normalizes it with the conversion-based form, sorts it with the same register network
// (one sample per register) and runs the complete sigma clipping of ssc_column.cuh on a
// shared-memory copy.  (The general kernel's list mode did this with a serial shared-memory bitonic
// sort: ~100 us of latency per launch whatever the list length.)
// ---------------------------------------------------------------------------------------------
template <int NP>
__global__ void __launch_bounds__(128)
k_sigma_handback(StackParams p, LeftoverList list) {
	__shared__ WORD col[NP * 128];
	__shared__ double s_unbias[65], s_inv_n[65], s_inv_n2[65];
	__shared__ unsigned long long blk_counters[SSC_COUNTERS];
	const int tid = threadIdx.x;
	unsigned int n_items = *list.count;
	if (n_items > list.capacity) n_items = list.capacity;
	if (blockIdx.x * 128u >= n_items) return;	// (usually the whole grid: the launch then costs its latency only)
	if (tid < SSC_COUNTERS) blk_counters[tid] = 0;
	if (tid >= 1 && tid <= 64) {
		s_unbias[tid] = tid > 1 ? ddiv((double) tid, (double) (tid - 1)) : 0.0;
		s_inv_n[tid] = ddiv(1.0, (double) tid);
		s_inv_n2[tid] = ddiv(1.0, (double) (tid * tid));
	}
	const SmallN tables{s_unbias, s_inv_n, s_inv_n2};
	__syncthreads();
	const ssc_band &b = p.band;
	const int W = b.width;
	const bool norm = p.coeff != nullptr && p.normalize != NO_NORM;
	for (unsigned int base = blockIdx.x * 128; base < n_items; base += gridDim.x * 128) {
		const unsigned int i = base + tid;
		if (i < n_items) {
			const unsigned int pix = list.ids[i];
			const unsigned int row = pix / (unsigned int) W;
			const int x = (int) (pix - row * (unsigned int) W);
			const int layer = (int) (row / (unsigned int) b.out_rows);
			const int rr = (int) (row - (unsigned int) layer * (unsigned int) b.out_rows);
			const int R = b.out_row0 + rr;
			const int srow = b.row_order == SSC_ROWS_SER ? b.height - 1 - R : R;
			const WORD *src = b.frames + (size_t) layer * b.layer_stride + (size_t) (srow - b.buf_row0) * W + x;
			unsigned int v[NP];
#pragma unroll
			for (int f = 0; f < NP; f++) v[f] = __ldg(src + (size_t) f * b.frame_stride);
			unsigned long long S = 0, Q = 0;
			if (norm) {
#pragma unroll
				for (int f = 0; f < NP; f++)
					v[f] = normalize_sample(v[f], p.normalize, p.coeff[f], p.coeff[NP + f], p.coeff[2 * NP + f]);
			}
#pragma unroll
			for (int f = 0; f < NP; f++) { S += v[f]; Q += sq(v[f]); }
			sort_network<NP>(v);		// packed min/max on values below 2^16: the upper halves stay 0
			Column c{col + tid, 128};
#pragma unroll
			for (int f = 0; f < NP; f++) c.set(f, v[f]);
			RejectState ist{0, 0, 0, 0};
			const unsigned int r = reject_sorted_and_mean<true>(c, NP, SIGMA, p.sig_low, p.sig_high, S, Q, ist, &tables);
			p.out[(size_t) layer * p.out_layer_stride + (size_t) (R - p.out_buf_row0) * W + x] = (WORD) r;
			if (ist.rej_lo) atomicAdd(&blk_counters[layer * 2 + 0], (unsigned long long) ist.rej_lo);
			if (ist.rej_hi) atomicAdd(&blk_counters[layer * 2 + 1], (unsigned long long) ist.rej_hi);
			if (ist.guard) atomicAdd(&blk_counters[6], 1ull);
			if (ist.cap) atomicAdd(&blk_counters[7], 1ull);
		}
	}
	__syncthreads();
	if (p.counters && tid < SSC_COUNTERS && blk_counters[tid])
		atomicAdd(&p.counters[tid], blk_counters[tid]);
}

// The register-network kernel applies when two adjacent pixels can be fetched as one
// aligned 32-bit word from every frame and nothing shifts the frames against each other.
bool fast_path_eligible(const StackParams &p, int is_median) {
	const ssc_band &b = p.band;
	if (b.n_frames < 2 || b.n_frames > 64) return false;
	if (!is_median && (b.shiftx || b.shifty)) return false;
	if ((b.width & 1) || (b.frame_stride & 1) || (b.layer_stride & 1)) return false;
	if (((uintptr_t) b.frames & 3) || ((uintptr_t) p.out & 3) || (p.out_layer_stride & 1)) return false;
	if ((size_t) b.layers * p.out_layer_stride >= 0xFFFFFFF0ull) return false;	// 32-bit result indices
	if ((unsigned long long) b.frame_stride * sizeof(WORD) >= 0xFFFFFFF0ull) return false;	// 32-bit frame stride in bytes
	return true;
}

template <int NP, bool FULL, int MODE, bool ROWTILE = false>
static int launch_fast_cfg(const StackParams &p, cudaStream_t stream) {
	const ssc_band &b = p.band;
	const unsigned long long total = (unsigned long long) b.layers * b.out_rows * (b.width >> 1);
	if (total == 0) return SSC_OK;
	const size_t smem = MODE == MODE_SIGMA_REG
			? (size_t) NP * FAST_THREADS * sizeof(unsigned int)
				+ (size_t) (FAST_THREADS / 32) * FAST_QUEUE_WORDS * FAST_QUEUE_CAP * sizeof(unsigned int)
			: (size_t) NP * FAST_THREADS * sizeof(unsigned int);
	if (cudaFuncSetAttribute(k_stack_fast<NP, FULL, MODE, ROWTILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess)
		return SSC_ERR_CUDA;
	// ROWTILE: tiles of four warp tiles, each 32 consecutive pairs of one row
	const unsigned long long tiles = ROWTILE
			? (unsigned long long) b.layers * b.out_rows * (((b.width >> 1) + 31) / 32)
			: (total + FAST_THREADS - 1) / FAST_THREADS;
	if (tiles > 0x7FFFFFFFull) return SSC_ERR_ARG;
	unsigned long long blocks = ROWTILE ? (tiles + FAST_THREADS / 32 - 1) / (FAST_THREADS / 32) : tiles;
	{	// a bounded grid of looping blocks: 8 per resident slot for big launches, 4 for small ones (measured;
		// a block should keep a few dozen tiles to amortise its prologue).  SSC_FAST_WAVES overrides.
		int dev = 0, sms = 148, per_sm = 1;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stack_fast<NP, FULL, MODE, ROWTILE>, FAST_THREADS, smem) != cudaSuccess || per_sm < 1)
			per_sm = 1;
		static int forced = -1;
		if (forced < 0) { const char *e = getenv("SSC_FAST_WAVES"); forced = e ? atoi(e) : 0; if (forced < 0) forced = 0; }
		const unsigned long long slots = (unsigned long long) sms * per_sm;
		const int waves = ROWTILE ? 1 : (forced ? forced : (tiles / slots >= 256 ? 8 : 4));	// ROWTILE: the resident warps draw their tiles
		const unsigned long long cap = slots * waves;
		if (blocks > cap) blocks = cap;
	}
	LeftoverList list{nullptr, 0, nullptr, nullptr};
	if (MODE == MODE_SIGMA_REG) {
		int rc = leftover_ids(total * 2, list, stream);
		if (rc) return rc;
	}
	k_stack_fast<NP, FULL, MODE, ROWTILE><<<(unsigned int) blocks, FAST_THREADS, smem, stream>>>(p, list, (unsigned int) tiles);
	if (cudaGetLastError() != cudaSuccess) return SSC_ERR_CUDA;
	if constexpr (MODE == MODE_SIGMA_REG) {	// the pixels that left the candidate window
		int dev = 0, sms = 148;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		k_sigma_handback<NP><<<sms * 8, 128, 0, stream>>>(p, list);
		if (cudaGetLastError() != cudaSuccess) return SSC_ERR_CUDA;
	}
	return SSC_OK;
}

template <int NP>
static int launch_fast_np(const StackParams &p, int is_median, cudaStream_t stream) {
	const bool full = p.band.n_frames == NP;
	if (is_median)
		return full ? launch_fast_cfg<NP, true, MODE_MEDIAN>(p, stream) : launch_fast_cfg<NP, false, MODE_MEDIAN>(p, stream);
	if (p.rejection == NO_REJEC)
		return full ? launch_fast_cfg<NP, true, MODE_NOREJ>(p, stream) : launch_fast_cfg<NP, false, MODE_NOREJ>(p, stream);
	if constexpr (NP >= 32) {
		// (non-negative parameters: a sample above the median is then never "low", one below never "high")
		if (full && p.rejection == SIGMA && p.sig_low >= 0.0 && p.sig_high >= 0.0) {
			const ssc_band &b = p.band;	// rows on 16-byte boundaries: the 16-byte gather (ROWTILE)
			const bool rows16 = b.width % 8 == 0 && b.frame_stride % 8 == 0 && b.layer_stride % 8 == 0
					&& ((uintptr_t) b.frames & 15) == 0;
			return rows16 ? launch_fast_cfg<NP, true, MODE_SIGMA_REG, true>(p, stream)
					: launch_fast_cfg<NP, true, MODE_SIGMA_REG, false>(p, stream);
		}
	}
	return full ? launch_fast_cfg<NP, true, MODE_QUEUE>(p, stream) : launch_fast_cfg<NP, false, MODE_QUEUE>(p, stream);
}

int launch_stack_fast(const StackParams &p, int is_median, cudaStream_t stream) {
	const int n = p.band.n_frames;
	if (n <= 16) return launch_fast_np<16>(p, is_median, stream);
	if (n <= 32) return launch_fast_np<32>(p, is_median, stream);
	return launch_fast_np<64>(p, is_median, stream);
}

}  // namespace ssc
